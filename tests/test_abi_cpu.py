"""CPU: the C-ABI library builds, loads and exports every symbol include/ampc_b200.h declares;
host-side argument checking and the loud no-GPU failure (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from autompc_b200 import build, _abi
    build.build()
    return _abi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ampc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ampc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(lib._name)
    for s in syms:
        assert hasattr(raw, s), "missing export %s" % s
    from autompc_b200 import _abi
    assert sorted(_abi.EXPORTS) == syms      # the ctypes binding covers exactly the header


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.ampc_version()
    assert isinstance(lib.ampc_last_error(), bytes)


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of every ABI struct as gcc sees the header == the ctypes mirror."""
    import subprocess
    from autompc_b200 import _abi
    structs = {"ampc_mppi_cfg": _abi.MppiCfg, "ampc_ilqr_cfg": _abi.IlqrCfg, "ampc_mlp_desc": _abi.MlpDesc,
               "ampc_quad_cost": _abi.QuadCost}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ampc_b200.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_host_side_validation_without_gpu(lib):
    import torch
    from autompc_b200 import MPPI, IterativeLQR, B200MLP, MLPWeights
    from autompc_b200.problems import cartpole_problem
    system, task, w, x0 = cartpole_problem()
    model = B200MLP(system, w)
    with pytest.raises(ValueError):
        MPPI(system, task, model, noise="mt19937")
    with pytest.raises(ValueError):
        MPPI(system, task, model, precision="fp8")
    with pytest.raises(ValueError):      # horizon 1: mppi.py:123 indexes act_sequence[-2]
        MPPI(system, task, model, horizon=1, precision="fp32")
    with pytest.raises(ValueError):      # non-MLP model
        MPPI(system, task, object.__new__(type("ARX", (), {"state_dim": 4})), precision="fp32")
    with pytest.raises(NotImplementedError):
        MLPWeights(w.W, w.b, "gelu", w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, 4, 1)
    with pytest.raises(ValueError):
        MLPWeights(w.W[:-1], w.b[:-1], "relu", w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, 4, 1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            MPPI(system, task, model, horizon=20, num_path=64, precision="fp32")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            IterativeLQR(system, task, model, horizon=10)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.pred(x0, np.zeros(1))


def test_unbounded_controls_rejected(lib):
    """Reference precondition: ctrl_scale = umax, so +-inf bounds give NaN."""
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.plugin import Task
    from autompc_b200.problems import cartpole_problem
    system, task, w, _ = cartpole_problem()
    t2 = Task(system)
    t2.set_cost(task.get_cost())
    with pytest.raises((ValueError, RuntimeError)) as ei:
        MPPI(system, t2, B200MLP(system, w), precision="fp32")
    import torch
    if torch.cuda.is_available():
        assert isinstance(ei.value, ValueError)


def test_plugin_subclasses_reference_abcs_when_available():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree only exists in the build container")
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import ref_loader; ns = ref_loader.load()\n"
            "import autompc_b200\n"
            "assert issubclass(autompc_b200.MPPI, ns.Controller)\n"
            "assert issubclass(autompc_b200.IterativeLQR, ns.Controller)\n"
            "assert issubclass(autompc_b200.B200MLP, ns.Model)\n"
            "assert issubclass(autompc_b200.MPPIFactory, ns.ControllerFactory)\n"
            "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
