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
