"""Dev tool (library built with -DV3_DEV): time the stand-alone height-scan kernel at 1 Mi envs for
several CTA shapes / occupancies / what-if modes.  mode bits: 1 no stores, 2 no gathers, 4 no math."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from shifu_b200 import _native as nv
n = int(os.environ.get("N", 1 << 20))
hp, raw = bench.build_a1(n, 0, 1, "cuda:0")
lib = nv.load()
fn = lib.shifu_debug_scan_only
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
def t(mode, cps, thr, k=10):
    run = lambda: nv.check(fn(hp.ctx.handle, nv.ptr(hp.root_state), nv.ptr(hp.obs_buf), mode, cps, thr, nv.current_stream()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): run()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / k
for np_ in (2, 4, 8):
    for thr, cps in ((192, 1), (192, 2), (192, 3), (192, 4), (192, 5), (384, 3), (384, 4)):
        print(f"pairs/batch={np_} threads={thr} ctas/sm={cps} warps/sm={thr//32*cps}: " +
              "  ".join(f"{nm}={t(m | (np_ << 8),cps,thr):.4f}" for m, nm in ((0, "full"), (1, "nostore"), (2, "nogather"), (3, "mathonly"))), flush=True)
