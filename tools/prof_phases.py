"""Dev tool: per-phase cycle counts of the pipelined fused kernel.
Build the library with SHIFU_NVCC_EXTRA=-DV3_PROFILE first (never ship that build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shifu_b200 import _native as nv

NAMES = {19: "TMA issue->w1 sees full_in", 0: "B.w0 loop top", 1: "B.w0 wait full_in(+s_free)", 2: "B.w0 B1", 3: "B.w0 barrier2", 13: "B.w0 B2",
         4: "B.w1 pre", 5: "B.w1 wait full_in", 6: "B.w1 B1", 7: "B.w1 barrier2", 14: "B.w1 B2(idle)",
         9: "scan head: fence + arrive", 21: "scan head: (env,dof) items", 22: "scan head: cla items", 23: "scan head: carry", 10: "DMA wait h_done", 11: "DMA store+reload",
         16: "scan wait e_done", 20: "scan prologue (item 0 index+gather)", 8: "scan wait b_done", 17: "scan work"}
COUNTS = {15: "B tiles", 12: "head tiles", 18: "scan tiles"}

def main():
    n = int(os.environ.get("N", 1 << 20))
    hp, raw = bench.build_a1(n, 0, 1, "cuda:0")
    lib = nv.load()
    buf = (ctypes.c_ulonglong * 32)()
    for _ in range(3):
        hp.step_resident(raw)
    lib.shifu_debug_profile(buf, 1)
    steps = 10
    for _ in range(steps):
        hp.post_physics()
    lib.shifu_debug_profile(buf, 1)
    v = list(buf)
    tiles = {k: v[k] for k in COUNTS}
    print("counts", {COUNTS[k]: tiles[k] for k in COUNTS})
    for k in sorted(NAMES):
        cnt = tiles[15] if k < 8 or k in (13, 14, 19) and k not in (8, 20) else (tiles[12] if k in (10, 11) else tiles[18])
        print(f"{NAMES[k]:32s} {v[k] / max(cnt, 1):10.0f} cycles/tile")

if __name__ == "__main__":
    main()
