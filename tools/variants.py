"""Dev tool: build library variants with extra nvcc -D flags and time the fused kernel for each.
  python tools/variants.py build name1="-DFOO=1" name2="-DBAR"     (container)
  python tools/variants.py run [steps]                               (GPU box)"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "tools", "_variants")
sys.path.insert(0, ROOT)

def build(specs):
    from shifu_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        os.remove(os.path.join(VDIR, f))
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags.split(), "-I", b.INCLUDE, "-o", out] + \
              [os.path.join(b.CSRC, s) for s in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "BUILD FAILED\n", r.stderr[-3000:]); continue
        spill = [l.strip() for l in (r.stdout + r.stderr).splitlines() if "spill" in l and "tma" not in l][:0]
        print("built", name, flags)

def run(steps):
    for f in sorted(os.listdir(VDIR)):
        env = dict(os.environ, SHIFU_B200_LIB=os.path.join(VDIR, f))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--no-cpu", "--steps", str(steps),
                            "--warmup", "5"], capture_output=True, text=True, env=env, timeout=300)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{f:40s} kernel_ms={j['roofline']['kernel_ms']:.4f} ms_per_step={j['ms_per_step']:.4f}", flush=True)
        except Exception as e:
            print(f, "FAILED", r.stdout[-300:], r.stderr[-800:], flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 30)
