"""Prints the tcgen05 kernel's timeline of CTA 0 for two horizon steps (debug; AMPC_TC_TRACE=1)."""
import ctypes as C, os, sys
import numpy as np
os.environ["AMPC_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autompc_b200 import MPPI, B200MLP, _abi
from autompc_b200.problems import halfcheetah_dim_problem
system, task, w, x0 = halfcheetah_dim_problem()
ctl = MPPI(system, task, B200MLP(system, w), horizon=50, num_path=16384, precision="bf16")
for _ in range(3): ctl.solve(x0)
EV = 120   # TRACE_EV in mppi_tc_kernel.cuh
buf = np.zeros(13 * EV, dtype=np.uint64)
n = _abi.lib().ampc_mppi_debug_trace(ctl._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size)
ev = []
for wp in range(13):
    for e in buf[wp * EV:(wp + 1) * EV]:
        if e: ev.append((int(e) >> 8, wp, int(e) & 255))
ev.sort()
t0 = 0   # cycles since kernel entry (CTA 0)
phase = {1: "setup done", 2: "left the horizon loop", 3: "CTA softmax record written", 4: "ticket / merge done", 5: "kernel end"}
names = {1: "MMA  saw bar_a", 2: "MMA  commit   ", 3: "EPI  saw bar_d", 4: "EPI  released ", 5: "EPI  next input",
         6: "EPI    ld#1 done (half = idx)", 7: "EPI    ld#2 done (half = idx)", 8: "EPI    stores issued (half = idx)",
         6: "MMA  saw bar_y (output layer complete), idx =", 9: "OWN    y loaded(0) / integrated(1) / input stored(2): idx ="}
for t, wp, tag in ev:
    if wp in (4, 8, 12):
        k, idx = tag >> 4, tag & 15
        if k == 11:
            print("%8d  warp %d  COST   %s" % (t - t0, wp, {0: "owner: state copy stored", 1: "helper: state copy visible", 2: "helper: stage cost done"}.get(idx, "?")))
        elif k == 10:
            print("%8d  warp %d  PHASE  %s" % (t - t0, wp, phase.get(idx, "?")))
        elif k >= 6:
            print("%8d  warp %d  %s %d" % (t - t0, wp, names.get(k, "?"), idx))
        else:
            print("%8d  warp %d  %s layer %d half/pair %d" % (t - t0, wp, names.get(k, "?"), idx >> 1, idx & 1))

# spread of the epilogue warps of CTA 0 (4-11): when each saw bar_d / released a half
print()
print("# per event: min / max over epilogue warps 4..11 (cycles since kernel entry), spread")
from collections import defaultdict
grp = defaultdict(list)
cnt = defaultdict(int)
for t, wp, tag in ev:
    if 4 <= wp <= 11 and (tag >> 4) in (3, 4, 5):
        key = (tag, cnt[(wp, tag)])
        cnt[(wp, tag)] += 1
        grp[key].append((t, wp))
for (tag, occ), lst in sorted(grp.items(), key=lambda kv: min(kv[1])):
    ts = [x[0] for x in lst]
    k, idx = tag >> 4, tag & 15
    print("%8d .. %8d  (+%4d, last = warp %d)  %s layer %d half %d" % (min(ts), max(ts), max(ts) - min(ts), max(lst)[1],
          names.get(k, "?"), idx >> 1, idx & 1))
