"""CPU tests (-m "not gpu"): the oracle's pins.

* Philox4x32-10 against the Random123 known-answer vectors;
* the torch-CPU restatement (oracle/shifu_oracle.py) against the golden fixtures that
  oracle/make_golden.py recorded from the UNMODIFIED reference (bit-exact, every tensor);
* in the build container (where /root/reference exists) the restatement against a fresh run of
  the unmodified reference on new seeds.
"""
import numpy as np
import pytest
import torch

from oracle import philox_np as px
from oracle import shifu_oracle as so
from tests import util


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in kat:
        out = px.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *key)
        assert " ".join(f"{int(v[0]):08x}" for v in out) == want


def test_philox_uniform_grid():
    u32 = np.array([0, 255, 256, 0xFFFFFFFF], dtype=np.uint32)
    u = px.u01_f32(u32)
    assert u[0] == 0 and u[1] == 0 and u[2] == np.float32(2.0 ** -24) and u[3] < 1.0
    assert px.randint10(u32, 10).tolist() == [0, 0, 0, 9]


def _exact(tag, got, want, skip=()):
    for k, w in want.items():
        if k in skip or k not in got:
            continue
        util.assert_close(f"{tag}:{k}", got[k], w, exact=True)


def _replay_a1(name, regenerate_map):
    z, meta = util.load_golden(name)
    n = meta["n"]
    if regenerate_map:
        from shifu_b200.sim import fake_isaacgym
        fake_isaacgym.install("cpu")
        from shifu_b200.configs import TerrainEnvConfig
        from shifu_b200.utils.heightmap import Terrain
        import hashlib
        np.random.seed(meta["map_seed"])
        ter = Terrain(TerrainEnvConfig().terrain, n)
        assert hashlib.sha256(ter.heightsamples.tobytes()).hexdigest() == meta["map_sha256"]
        hs = ter.heightsamples
    else:
        hs = z["height_samples"]
    p, st = util.make_oracle_a1(n, hs, z["terrain_origins"], z["terrain_types"], z["env_origins_init"],
                                border_size=int(meta["border_size"]), max_terrain_level=meta["max_terrain_level"],
                                num_cols=meta["num_cols"], rng_seed=meta["rng_seed"])
    snap = util.golden_snap(z, 0)
    so.a1_reset(p, st, snap)
    _exact(f"{name}/reset", util.oracle_a1_outputs(st), util.golden_out(z, 0))
    assert np.array_equal(st.reset_ids.numpy(), z["s0/reset_ids"])
    st.ep_len[:] = torch.from_numpy(z["ep_len_init"])
    st.terrain_levels[:] = torch.from_numpy(z["levels_init"])
    for t in range(1, meta["steps"] + 1):
        snap = util.golden_snap(z, t)
        so.a1_step(p, st, snap.actions, snap)
        _exact(f"{name}/s{t}", util.oracle_a1_outputs(st), util.golden_out(z, t))
        assert np.array_equal(st.reset_ids.numpy(), z[f"s{t}/reset_ids"])
    return z, meta


def test_oracle_reproduces_golden_a1_small():
    z, meta = _replay_a1("a1_small", regenerate_map=False)
    # the crafted zero-reset step really had no resets and the extras persisted
    t = meta["zero_reset_step"]
    assert z[f"s{t}/out/reset"].sum() == 0
    for k in so.A1_REWARD_TERMS:
        assert z[f"s{t}/out/extras/{k}"] == z[f"s{t - 1}/out/extras/{k}"]


def test_oracle_reproduces_golden_a1_fullmap():
    _replay_a1("a1_fullmap", regenerate_map=True)


def test_golden_microcases_a1_small():
    """Hand-checkable rows of the fixture."""
    z, meta = util.load_golden("a1_small")
    o1 = util.golden_out(z, 1)
    # envs 6, 7: base contact force of norm exactly 1.0 -> NOT a contact termination ('> 1.')
    assert o1["contact_term"][6] == 0 and o1["contact_term"][7] == 0
    # envs 8, 9, 10 start at ep_len 499/500/501 -> after +1: 500 (no), 501 (time-out), 502 (time-out)
    assert o1["time_out"][8] == 0 and o1["time_out"][9] == 1 and o1["time_out"][10] == 1
    # env 4 has the identity quaternion: the scan grid is axis aligned, so the cell indices are
    # trunc((p + x + border)/0.1) and the heights follow from the stored map
    hs = z["height_samples"]
    root = z["s1/in/root_offset"][4].copy()
    env_origin = util.golden_out(z, 0)["env_origins"][4]
    pos = (root[:3] + env_origin).astype(np.float32)
    xs = np.array(so.A1Params(n=1).points_x, dtype=np.float32)
    ys = np.array(so.A1Params(n=1).points_y, dtype=np.float32)
    want = np.zeros(187, dtype=np.float32)
    for k in range(187):
        fx = np.float32(np.float32(np.float32(xs[k % 17] + pos[0]) + np.float32(5)) / np.float32(0.1))
        fy = np.float32(np.float32(np.float32(ys[k // 17] + pos[1]) + np.float32(5)) / np.float32(0.1))
        ix = int(np.clip(int(fx), 0, hs.shape[0] - 2))
        iy = int(np.clip(int(fy), 0, hs.shape[1] - 2))
        want[k] = np.float32(min(hs[ix, iy], hs[ix + 1, iy], hs[ix, iy + 1])) * np.float32(0.005)
    assert np.array_equal(o1["measured_heights"][4], want)
    # off-map robots (envs 1-3) clamp to the map border cells
    assert np.all(o1["measured_heights"][1] == np.float32(min(hs[0, 0], hs[1, 0], hs[0, 1])) * np.float32(0.005))


def test_oracle_reproduces_golden_abb():
    z, meta = util.load_golden("abb_small")
    p = so.AbbParams(n=meta["n"], rng_seed=meta["rng_seed"])
    st = so.abb_new_state(p)
    st.ep_len[:] = torch.from_numpy(z["ep_len_init"])
    for t in range(1, meta["steps"] + 1):
        so.abb_step(p, st, util.golden_snap(z, t, "abb"))
        got = dict(obs=st.obs, rew=st.rew, reset=st.reset.to(torch.uint8), time_out=st.time_out.to(torch.uint8),
                   success=st.success.to(torch.uint8), ep_len=st.ep_len, root_state=st.root_state,
                   dof_state=st.dof_state)
        for k in so.ABB_REWARD_TERMS:
            got["ep_sum/" + k] = st.ep_sums[k]
        for k, v in st.extras["episode"].items():
            got["extras/" + k] = torch.as_tensor(v)
        got = {k: v.numpy() for k, v in got.items()}
        _exact(f"abb/s{t}", got, util.golden_out(z, t), skip=("dof_targets",))
        assert np.array_equal(st.reset_ids.numpy(), z[f"s{t}/reset_ids"])


@pytest.mark.reference
def test_oracle_matches_unmodified_reference_fresh_seeds():
    """Build container only: drive /root/reference itself on new seeds and compare bit for bit."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from shifu_b200.sim.synthetic import a1_snapshot, abb_snapshot
    ns = rh.load_reference()
    n, seed = 96, 2024
    env = rh.make_a1(ns, n)
    isg = env.isg_env
    p, st = util.make_oracle_a1(n, isg.height_samples.clone(), isg.terrain_origins.clone(),
                                isg.terrain_types.clone(), isg.env_origins.clone())
    ep = np.random.RandomState(5).randint(0, 500, size=n)
    rec, _, rid = rh.run_a1(ns, env, seed=seed, steps=3, ep_len_init=ep, snap_kw=dict(p_base=0.08))
    so.a1_reset(p, st, a1_snapshot(seed, 0, n, p_base=0.08))
    _exact("ref/reset", util.oracle_a1_outputs(st), rec[0])
    st.ep_len[:] = torch.from_numpy(ep)
    for t in range(1, 4):
        snap = a1_snapshot(seed, t, n, p_base=0.08)
        so.a1_step(p, st, snap.actions, snap)
        _exact(f"ref/s{t}", util.oracle_a1_outputs(st), rec[t])
        assert np.array_equal(st.reset_ids.numpy(), rid[t][-1])
    # ABB
    n = 64
    env = rh.make_abb(ns, n)
    ep = np.random.RandomState(6).randint(0, 200, size=n)
    rec, _, rid = rh.run_abb(ns, env, seed=seed, steps=3, ep_len_init=ep)
    pa = so.AbbParams(n=n)
    sa = so.abb_new_state(pa)
    sa.ep_len[:] = torch.from_numpy(ep)
    for t in range(1, 4):
        so.abb_step(pa, sa, abb_snapshot(seed, t, n))
        assert np.array_equal(sa.obs.numpy(), rec[t - 1]["obs"])
        assert np.array_equal(sa.rew.numpy(), rec[t - 1]["rew"])
        assert np.array_equal(sa.reset.numpy().astype(np.uint8), rec[t - 1]["reset"])
        assert np.array_equal(sa.root_state.numpy(), rec[t - 1]["root_state"])


def test_oracle_arm_ik_matches_reference_fixture():
    """Row N2: the restated arm action path vs outputs recorded from the unmodified reference
    (AbbRobot.step of the env, and shifu.utils.torch_utils.inverse_kinematics).  torch.inverse is
    LAPACK: identical on the machine that made the fixture, 1e-6 elsewhere."""
    import json
    import os
    from oracle import shifu_oracle as so
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "arm_ik.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    d = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    goal = so.arm_goal_from_actions(d["ee_pose"][:, :3], d["actions"], meta["ee_velocity"], meta["dt"],
                                    meta["min_ee_pos"], meta["max_ee_pos"], meta["tar_quat"])
    got = so.arm_ik(d["dof_pos"], d["ee_pose"], d["j_ee"], goal, meta["damping"]).numpy()
    np.testing.assert_allclose(got, z["out/dof_targets_step"], rtol=1e-6, atol=1e-6)
    got = so.arm_ik(d["dof_pos"], d["ee_pose"], d["j_ee"], d["goal_pose"], meta["damping"]).numpy()
    np.testing.assert_allclose(got, z["out/dof_targets_goal"], rtol=1e-6, atol=1e-6)
    # the clamp is active for some envs and inactive for others
    raw = d["ee_pose"][:, :3] + d["actions"] * meta["ee_velocity"] * meta["dt"]
    clamped = (goal[:, :3] != raw).any(dim=1)
    assert 0 < int(clamped.sum()) < len(clamped)


def test_oracle_camera_refresh_matches_reference_fixture():
    """Row N4: restated CameraSensor.refresh_image_tensors vs buffers recorded from the unmodified
    reference class driven through IsaacGymEnv.refresh_sensors (bit-exact, -0.0 and inf included)."""
    import json
    import os
    from oracle import shifu_oracle as so
    from shifu_b200.sim.fake_isaacgym import synthetic_camera_image
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    keys = ("color", "depth", "seg", "flow")
    ins = {k: [torch.from_numpy(x) for x in z["in/" + k]] for k in keys}
    for t, k in enumerate(keys):          # the stand-in renderer still produces the recorded frames
        assert torch.equal(ins[k][1], synthetic_camera_image(1, t, meta["frame"], meta["height"], meta["width"]))
    for norm, tag in ((False, "raw"), (True, "norm")):
        got = so.camera_refresh(**ins, image_normalization=norm)
        for k in keys:
            want = z[f"{tag}/{k}"]
            assert got[k].numpy().dtype == want.dtype
            assert got[k].numpy().tobytes() == want.tobytes(), (k, tag)
    assert np.signbit(z["norm/depth"][0, 0, 0]) and z["norm/depth"][0, 0, 0] == 0.0
