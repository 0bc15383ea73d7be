"""Builds libampc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libampc_b200.so")
SOURCES = ["mppi_api.cu", "mppi_fp32.cu", "mppi_tc.cu", "mlp_ops.cu", "ilqr.cu", "linear.cu"]
# the tcgen05 kernel template is instantiated once per (cta_group, NXP, ReLU, trace) combination, each as its own
# compilation of mppi_tc_inst.cu, so that the matrix builds in parallel
# ... and, for ReLU networks, in "dz" mode (see mppi_tc_kernel.cuh); the last two are the timeline builds
TC_INSTANCES = [(cg, nxp, relu, f16, 0, dz) for dz in (0, 1) for f16 in (0, 1) for cg in (1, 2) for relu in (0, 1)
                for nxp in (4, 8, 16, 24, 32) if relu or not dz] + [(2, 24, 1, 0, 1, 0), (2, 24, 1, 0, 1, 1)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "ampc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_dir():
    """Objects are cached OUTSIDE the repo (only the .so travels with the snapshot): a source whose object is newer
    than the source and every header is not recompiled."""
    d = os.environ.get("AMPC_OBJ_DIR", os.path.join("/tmp", "ampc_b200_obj"))
    os.makedirs(d, exist_ok=True)
    return d


def _fresh(obj, src):
    if not os.path.exists(obj):
        return False
    t = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    deps += [os.path.join(os.path.dirname(HERE), "include", "ampc_b200.h"), os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a (up to os.cpu_count() nvcc processes at a time) and link the library."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    odir = _obj_dir()
    jobs = [(src, os.path.join(odir, src.replace(".cu", ".o")), []) for src in SOURCES]
    for cg, nxp, relu, f16, tr, dz in TC_INSTANCES:
        obj = os.path.join(odir, "mppi_tc_inst_cg%d_nxp%d_relu%d_f16%d_trace%d_dz%d.o" % (cg, nxp, relu, f16, tr, dz))
        jobs.append(("mppi_tc_inst.cu", obj, ["-DAMPC_TC_INST_CG=%d" % cg, "-DAMPC_TC_INST_NXP=%d" % nxp,
                                              "-DAMPC_TC_INST_RELU=%d" % relu, "-DAMPC_TC_INST_F16=%d" % f16,
                                              "-DAMPC_TC_INST_TRACE=%d" % tr, "-DAMPC_TC_INST_DZ=%d" % dz]))
    max_par = max(1, min(len(jobs), int(os.environ.get("AMPC_BUILD_JOBS", os.cpu_count() or 4))))
    objs = [obj for _, obj, _ in jobs]
    pending, running = [j for j in jobs if force or not _fresh(j[1], j[0])], []
    # development knob: AMPC_TC_DEV_ONLY="cg2_nxp24_relu1" rebuilds only the matching kernel instantiations and links
    # the others from their (stale) cached objects -- never for a build that is tested or shipped
    dev = os.environ.get("AMPC_TC_DEV_ONLY")
    if dev:
        pending = [j for j in pending if j[0] != "mppi_tc_inst.cu" or dev in j[1] or not os.path.exists(j[1])]

    def reap(src, pr):
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            for _, other in running:
                other.kill()
            raise RuntimeError("nvcc failed on %s" % src)

    while pending or running:
        while pending and len(running) < max_par:
            src, obj, defs = pending.pop(0)
            cmd = [_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + [
                "-c", os.path.join(CSRC, src), "-o", obj]
            running.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        reap(*running.pop(0))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
