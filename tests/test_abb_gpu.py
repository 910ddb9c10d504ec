"""GPU parity tests of the ABB push-box prior-stage kernel (row a16), through the C-ABI."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _make(n, rng_seed=0x5EED):
    from shifu_b200 import hotpath
    dev = torch.device("cuda:0")
    root = torch.zeros(n * 4, 13, device=dev)
    root[:, 6] = 1
    body = torch.zeros(n * 10, 13, device=dev)
    body[:, 6] = 1
    dof = torch.zeros(n * 6, 2, device=dev)
    return hotpath.AbbHotPath(hotpath.abb_desc(n, rng_seed=rng_seed), root_state=root, body_state=body,
                              dof_state=dof)


def _step(hp, snap):
    n = hp.n
    hp.root_state.view(n, 4, 13).copy_(snap.root.cuda())
    hp.body_state.view(n, 10, 13).copy_(snap.body.cuda())
    hp.dof_state.view(n, 6, 2).copy_(snap.dof.cuda())
    hp.post_physics()
    hp.finalize()


def _outputs(hp):
    d = dict(obs=hp.obs_buf, rew=hp.rew_buf, reset=hp.reset_buf.to(torch.uint8),
             time_out=hp.time_out_buf.to(torch.uint8), success=hp.success_buf.to(torch.uint8),
             ep_len=hp.ep_len, root_state=hp.root_state, dof_state=hp.dof_state)
    for k in hp.terms:
        d["ep_sum/" + k] = hp.ep_sums[k]
    for k, v in hp.extras()["episode"].items():
        d["extras/" + k] = v
    out = {k: v.detach().cpu().numpy().copy() for k, v in d.items()}
    out["reset_ids"] = hp.reset_id_list().cpu().numpy().copy()
    return out


def test_golden_abb():
    z, meta = util.load_golden("abb_small")
    hp = _make(meta["n"], meta["rng_seed"])
    hp.ep_len.copy_(torch.from_numpy(z["ep_len_init"]))
    for t in range(1, meta["steps"] + 1):
        _step(hp, util.golden_snap(z, t, "abb"))
        got = _outputs(hp)
        util.compare_a1(got, util.golden_out(z, t), f"abb/s{t}", skip=("dof_targets",))
        assert np.array_equal(got["reset_ids"], z[f"s{t}/reset_ids"])


@pytest.mark.parametrize("n", [1, 77, 65536])
def test_oracle_parity_abb(n):
    """BASELINE config 4 size (65 536 envs) included."""
    from oracle import shifu_oracle as so
    from shifu_b200.sim.synthetic import abb_snapshot
    p = so.AbbParams(n=n)
    st = so.abb_new_state(p)
    hp = _make(n)
    ep = torch.from_numpy(np.random.RandomState(n).randint(0, 200, size=n))
    st.ep_len[:] = ep
    hp.ep_len.copy_(ep)
    for t in range(1, 4):
        snap = abb_snapshot(21, t, n)
        so.abb_step(p, st, snap)
        _step(hp, snap)
        got = _outputs(hp)
        want = dict(obs=st.obs, rew=st.rew, reset=st.reset.to(torch.uint8), time_out=st.time_out.to(torch.uint8),
                    success=st.success.to(torch.uint8), ep_len=st.ep_len, root_state=st.root_state,
                    dof_state=st.dof_state)
        for k in so.ABB_REWARD_TERMS:
            want["ep_sum/" + k] = st.ep_sums[k]
        if "episode" in st.extras:
            for k, v in st.extras["episode"].items():
                want["extras/" + k] = torch.as_tensor(v)
        want = {k: v.numpy() for k, v in want.items()}
        util.compare_a1(got, want, f"abb n{n}/s{t}")
        assert np.array_equal(got["reset_ids"], st.reset_ids.numpy())


# ---------------------------------------------------------------------------------------------
# Row N2: shifu_arm_ik — floating-point kernel (6x6 damped least squares).  Tolerance: the CUDA
# result must be as close to exact (float64) arithmetic as the reference's own float32 path is,
# up to a factor 2, and within 3x that distance (+1e-6) of the reference's float32 output.
# ---------------------------------------------------------------------------------------------
def _arm_ik_cuda(d, meta, mode):
    from shifu_b200 import hotpath
    n = d["j_ee"].shape[0]
    k = hotpath.EnvKernels("cuda:0", n)
    bodies, ee_body, links, ee_link = 10, 6, 9, 5
    body = torch.zeros(n * bodies, 13, device="cuda")
    body.view(n, bodies, 13)[:, ee_body, :7] = d["ee_pose"].cuda()
    jac = torch.randn(n, links, 6, 6, device="cuda")          # other links hold garbage on purpose
    jac[:, ee_link] = d["j_ee"].cuda()
    dof = torch.zeros(n * 6, 2, device="cuda")
    dof.view(n, 6, 2)[:, :, 0] = d["dof_pos"].cuda()
    dof.view(n, 6, 2)[:, :, 1] = 7.0
    out = torch.full((n, 6), float("nan"), device="cuda")
    kw = dict(body_state=body, num_bodies=bodies, ee_body=ee_body, jacobian=jac, ee_link=ee_link, dof_state=dof,
              num_dof=6, dof_targets=out, damping=meta["damping"])
    if mode == "step":
        k.arm_ik(actions=d["actions"].cuda(), ee_velocity=meta["ee_velocity"], dt=meta["dt"],
                 min_ee_pos=meta["min_ee_pos"], max_ee_pos=meta["max_ee_pos"], tar_quat=meta["tar_quat"], **kw)
    else:
        k.arm_ik(goal_pose=d["goal_pose"].cuda(), **kw)
    return out.cpu()


def _arm_ik_check(d, meta, mode, ref32=None):
    from oracle import shifu_oracle as so
    goal = (so.arm_goal_from_actions(d["ee_pose"][:, :3], d["actions"], meta["ee_velocity"], meta["dt"],
                                     meta["min_ee_pos"], meta["max_ee_pos"], meta["tar_quat"])
            if mode == "step" else d["goal_pose"])
    o32 = so.arm_ik(d["dof_pos"], d["ee_pose"], d["j_ee"], goal, meta["damping"])
    o64 = so.arm_ik(d["dof_pos"].double(), d["ee_pose"].double(), d["j_ee"].double(), goal.double(), meta["damping"])
    got = _arm_ik_cuda(d, meta, mode)
    assert torch.isfinite(got).all()
    err_ref = float((o32.double() - o64).abs().max())
    err_cuda = float((got.double() - o64).abs().max())
    # the reference inverts J J^T + l^2 I with LAPACK (torch.inverse) in fp32, the kernel solves by Cholesky:
    # both are fp32 roundings of the float64 answer, so the bar is "no farther from float64 than the
    # reference itself" (x2), and the achieved numbers are printed for the record (pytest -s / -rP)
    err_vs_ref = float((got - o32).abs().max())
    rel = float(((got.double() - o64).abs() / o64.abs().clamp_min(1e-3)).max())
    print(f"arm_ik[{mode}]: |cuda - f64| max {err_cuda:.3e} (reference fp32: {err_ref:.3e}), |cuda - ref32| max "
          f"{err_vs_ref:.3e}, max rel vs f64 {rel:.3e}")
    assert err_cuda <= max(2.0 * err_ref, 1e-6), (mode, err_cuda, err_ref)
    assert err_vs_ref <= 3.0 * err_ref + 1e-6
    if ref32 is not None:
        assert float((got - torch.from_numpy(ref32)).abs().max()) <= 3.0 * err_ref + 1e-6


@pytest.mark.parametrize("mode", ["step", "goal"])
def test_arm_ik_matches_reference_fixture(mode):
    import json, os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "arm_ik.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    d = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    _arm_ik_check(d, meta, mode, ref32=z[f"out/dof_targets_{mode}"])


@pytest.mark.parametrize("n", [1, 4099])
def test_arm_ik_random(n):
    from oracle.make_golden import arm_ik_inputs
    meta = dict(ee_velocity=0.2, dt=0.1, min_ee_pos=[-0.2, -0.2, 0.11], max_ee_pos=[0.2, 0.2, 0.14],
                tar_quat=[0., 1., 0., 0.], damping=0.05)
    d = arm_ik_inputs(100 + n, n)
    _arm_ik_check(d, meta, "step")
    _arm_ik_check(d, meta, "goal")


def test_arm_ik_argument_errors():
    from shifu_b200 import hotpath, _native as nv
    k = hotpath.EnvKernels("cuda:0", 4)
    z = lambda *s: torch.zeros(*s, device="cuda")
    kw = dict(body_state=z(4 * 2, 13), num_bodies=2, ee_body=1, jacobian=z(4, 1, 6, 6), ee_link=0,
              dof_state=z(4 * 6, 2), num_dof=6, dof_targets=z(4, 6))
    with pytest.raises(nv.ShifuNativeError):
        k.arm_ik(**kw)                                        # neither goal_pose nor actions
    with pytest.raises(nv.ShifuNativeError):
        k.arm_ik(goal_pose=z(4, 7), **{**kw, "ee_body": 5})   # body index out of range
