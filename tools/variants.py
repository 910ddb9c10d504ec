"""Dev tool: build library variants with extra nvcc -D flags and time the fused kernel for each.
  python tools/variants.py build name1="-DFOO=1" name2="-DBAR"     (container; parallel builds)
  python tools/variants.py run [steps]                               (GPU box)
A name starting with "prof_" is built with -DV3_PROFILE and run through tools/prof_phases.py."""
import os, subprocess, sys, json
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "tools", "_variants")
sys.path.insert(0, ROOT)

def build(specs):
    from shifu_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        os.remove(os.path.join(VDIR, f))
    def one(spec):
        name, _, flags = spec.partition("=")
        if name.startswith("prof_"):
            flags += " -DV3_PROFILE"
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags.split(), "-I", b.INCLUDE, "-o", out] + \
              [os.path.join(b.CSRC, s) for s in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return f"{name} BUILD FAILED\n{r.stderr[-3000:]}"
        regs = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l]
        log = (r.stdout + r.stderr)
        i = log.find("a1_post_physics_tma_kernelILb0ELb0")
        tail = log[i:i + 600].replace("\n", " ") if i >= 0 else ""
        import re
        m = re.search(r"(\d+) bytes spill stores.*?Used (\d+) registers", tail)
        return f"built {name} {flags} | tma<0,0>: spill {m.group(1) if m else '?'} regs {m.group(2) if m else '?'}"
    with ThreadPoolExecutor(max_workers=8) as ex:
        for msg in ex.map(one, specs):
            print(msg, flush=True)

def run(steps):
    for f in sorted(os.listdir(VDIR)):
        env = dict(os.environ, SHIFU_B200_LIB=os.path.join(VDIR, f))
        if f.startswith("lib_prof_"):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_phases.py")], capture_output=True,
                               text=True, env=env, timeout=300)
            print("==", f, "\n" + r.stdout[-2500:], r.stderr[-500:], flush=True)
            continue
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
