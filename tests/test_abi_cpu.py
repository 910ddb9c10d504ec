"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a without a GPU, loads,
exports every symbol include/shifu_b200.h declares, its ctypes mirror has the header's exact
struct layout, and argument errors surface as error codes (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shifu_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shifu_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from shifu_b200 import _native as nv, build
    path = build.build()
    assert os.path.exists(path)
    lib = nv.load()
    declared = _declared_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in nv.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(nv.SIGNATURES) == declared
    assert lib.shifu_abi_version() == nv.ABI_VERSION


def test_sm100a_cubin_present():
    """The .so must carry sm_100a SASS (not PTX for another arch)."""
    from shifu_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_ctypes_structs_match_header_layout(tmp_path):
    from shifu_b200 import _native as nv
    structs = {"ShifuA1Desc": nv.A1Desc, "ShifuA1StepIO": nv.A1StepIO, "ShifuAbbDesc": nv.AbbDesc,
               "ShifuAbbStepIO": nv.AbbStepIO, "ShifuArmIkIO": nv.ArmIkIO,
               "ShifuCameraGatherIO": nv.CameraGatherIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_argument_errors_without_gpu():
    from shifu_b200 import _native as nv, hotpath
    lib = nv.load()
    out = C.c_void_p()
    assert lib.shifu_ctx_create(0, None, None, C.byref(out)) == nv.E_NULL
    d = hotpath.a1_desc(0)
    assert lib.shifu_ctx_create(0, C.byref(d), None, C.byref(out)) == nv.E_RANGE
    assert b"num_envs" in lib.shifu_last_error()
    d = hotpath.a1_desc(16)
    d.num_dof = 7
    assert lib.shifu_ctx_create(0, C.byref(d), None, C.byref(out)) == nv.E_RANGE
    d = hotpath.a1_desc(16)
    d.abi_version = 99
    assert lib.shifu_ctx_create(0, C.byref(d), None, C.byref(out)) == nv.E_RANGE
    a = hotpath.abb_desc(16)
    a.ee_body = 99
    assert lib.shifu_ctx_create(0, None, C.byref(a), C.byref(out)) == nv.E_RANGE
    assert lib.shifu_ctx_create_util(0, 0, C.byref(out)) == nv.E_RANGE
    import torch
    if not torch.cuda.is_available():
        # a valid descriptor still cannot create a context: there is no CPU fallback
        assert lib.shifu_ctx_create(0, C.byref(hotpath.a1_desc(16)), None, C.byref(out)) == nv.E_NODEVICE
        with pytest.raises(nv.ShifuNativeError):
            hotpath.EnvKernels("cpu", 16)


def test_reward_term_registry():
    from shifu_b200 import _native as nv, hotpath
    comp = hotpath.compile_reward_terms(["torques_penalize", "tracking_lin_vel"], hotpath.A1_TERM_CODES,
                                        hotpath.A1_TERM_PARAMS)
    assert [c for c, _, _ in comp] == [nv.REW_TORQUES, nv.REW_TRACKING_LIN_VEL]          # list order kept
    assert comp[0][1] == pytest.approx(-2e-5) and comp[1][1:] == (1.0, 0.25)
    with pytest.raises(KeyError):
        hotpath.compile_reward_terms(["feet_air_time"], hotpath.A1_TERM_CODES, hotpath.A1_TERM_PARAMS)
    with pytest.raises(ValueError):
        hotpath.compile_reward_terms([], hotpath.A1_TERM_CODES, hotpath.A1_TERM_PARAMS)
    d = hotpath.a1_desc(8, terms=["leg_collision", "smoothing_action"])
    assert d.num_reward_terms == 2 and list(d.reward_terms)[:2] == [nv.REW_LEG_COLLISION, nv.REW_SMOOTHING_ACTION]
    assert d.level_up_distance == 4.0 and d.max_episode_length == 500 and list(d.leg_bodies) == [2, 3, 6, 7, 10, 11, 14, 15]


def test_exact_division_by_point_one(tmp_path):
    """csrc/exact_math.cuh::div_const == IEEE x/0.1f for every float with 1e-3 <= |x| <= 65536
    (the height-cell index chain, shifu/gym/isaac_gym.py:421).  Checked exhaustively in C with the
    same three operations (fmaf is the correctly rounded fused multiply-add)."""
    src = tmp_path / "div.c"
    src.write_text(r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void){ const float d=0.1f, r=1.0f/d; uint32_t a,b; float lo=1e-3f, hi=65536.f; unsigned long long bad=0,n=0;
  memcpy(&a,&lo,4); memcpy(&b,&hi,4);
  for(uint32_t u=a;u<=b;++u){ float x; memcpy(&x,&u,4);
    for(int s=0;s<2;++s){ float xx=s?-x:x; float q0=xx*r; float e=fmaf(-d,q0,xx); float q=fmaf(e,r,q0); if(q!=xx/d) bad++; n++; } }
  printf("%llu %llu\n", n, bad); return 0; }''')
    exe = tmp_path / "div"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    n, bad = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()
    assert int(n) > 400_000_000 and int(bad) == 0


def test_torch_philox_matches_oracle_philox():
    import numpy as np
    import torch
    from oracle import philox_np as px
    from shifu_b200.utils import philox
    ids = torch.tensor([0, 1, 5, 1 << 20, (1 << 31) + 7])
    for stream in (0, 1, 3, 7):
        want = px.draw_u32(0x5EED, ids.numpy(), 12345, stream)
        got = philox._lanes(0x5EED, ids, 12345, stream)
        for lane in range(4):
            assert np.array_equal(got[lane].numpy().astype(np.uint32), want[lane])
    u = philox.draw_u01(0x5EED, ids, 3, 1, 2).numpy()
    assert np.array_equal(u, px.u01_f32(px.draw_u32(0x5EED, ids.numpy(), 3, 1)[:2]).T)
    r = philox.draw_randint(0x5EED, ids, 3, 0, 10).numpy()
    assert np.array_equal(r, px.randint10(px.draw_u32(0x5EED, ids.numpy(), 3, 0)[0], 10))


def test_build_is_keyed_on_source_content_not_mtimes(tmp_path, monkeypatch):
    """One process per GPU shares the tree: a touched-but-unchanged source (git checkout, a
    snapshot copy) must not make every rank rebuild the library; a changed source must."""
    import shutil
    from shifu_b200 import build as b
    assert not b.needs_build()
    pkg = tmp_path / "shifu_b200"
    shutil.copytree(b.PKG_DIR, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.tmp*"))
    shutil.copytree(b.INCLUDE, tmp_path / "include")
    monkeypatch.setattr(b, "PKG_DIR", str(pkg))
    monkeypatch.setattr(b, "CSRC", str(pkg / "csrc"))
    monkeypatch.setattr(b, "INCLUDE", str(tmp_path / "include"))
    monkeypatch.setattr(b, "LIB_PATH", str(pkg / "libshifu_b200.so"))
    monkeypatch.setattr(b, "STAMP_PATH", str(pkg / "libshifu_b200.so.srchash"))
    assert not b.needs_build()                        # relocated copy: same content, same stamp
    src = pkg / "csrc" / "philox.cuh"
    os.utime(src)                                     # newer mtime, same bytes
    assert not b.needs_build()
    src.write_text(src.read_text() + "\n// changed\n")
    assert b.needs_build()
