"""dev helper: launch every kernel of libshifu_b200.so once or twice at its benchmark size, for
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        --csv --log-file gpurun_out/r2_kernels.csv python tools/ncu_all_kernels.py
(`tools/ncu_kernels_summary.py` turns the CSV into profiles/r2_kernels_ncu.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from shifu_b200 import hotpath
from shifu_b200.sim.synthetic import abb_snapshot

dev = "cuda:0"
n = int(os.environ.get("N", 1 << 20))
for store in (False, True):                      # both instantiations of the fused kernel
    env, raw = bench.build_env(n, 0, 1, dev, store_heights=store, use_graph=False)
    hp = env.hot
    for _ in range(3):
        env.step(raw)
    if store:
        hp.body_frame()
        hp.get_heights()
        hp.eval_terms()
        ids = hp.reset_id_list().clone()
        hp.reset_idx(ids)
        k = env.isg_env.kernels()
        k.clip(hp.obs_buf, 100.0)
        k.history_add(hp.history, hp.actions)
        k.nonzero(hp.reset_buf)
    torch.cuda.synchronize()
    del env, hp
    torch.cuda.empty_cache()
# ABB prior stage, 65 536 envs (BASELINE configs[3]) + arm IK + camera gather
m = 65536
snap = abb_snapshot(7, 1, m, gen_device=dev)
ab = hotpath.AbbHotPath(hotpath.abb_desc(m), root_state=snap.root.reshape(m * 4, 13).contiguous(),
                        body_state=snap.body.reshape(m * 10, 13).contiguous(), dof_state=snap.dof.reshape(m * 6, 2).contiguous())
ab.ep_len.copy_(torch.randint(0, 200, (m,), device=dev))
for _ in range(2):
    ab.step_resident()
ab.reset_idx(ab.reset_id_list().clone())
k = hotpath.EnvKernels(dev, m)
jac = (torch.rand(m, 9, 6, 6, device=dev) * 2 - 1) * 0.6
act = torch.rand(m, 3, device=dev) * 2 - 1
for _ in range(2):
    k.arm_ik(body_state=ab.body_state, num_bodies=10, ee_body=6, jacobian=jac, ee_link=5, dof_state=ab.dof_state, num_dof=6,
             dof_targets=ab.dof_targets, actions=act, ee_velocity=0.2, dt=0.1, min_ee_pos=(-0.2, -0.2, 0.11),
             max_ee_pos=(0.2, 0.2, 0.14), tar_quat=(0., 1., 0., 0.))
c, h, w = 2048, 128, 128
color = torch.randint(0, 256, (c, h, w, 4), dtype=torch.uint8, device=dev)
depth = -torch.rand(c, h, w, device=dev)
seg = torch.randint(0, 5, (c, h, w), dtype=torch.int32, device=dev)
table = lambda t: (t.data_ptr() + torch.arange(c, dtype=torch.int64) * t[0].numel() * t.element_size()).to(dev)
kc = hotpath.EnvKernels(dev, c)
for _ in range(2):
    kc.camera_gather(height=h, width=w, normalize_color=True, color=(table(color), torch.empty(c, h, w, 3, device=dev)),
                     depth=(table(depth), torch.empty(c, h, w, device=dev)),
                     seg=(table(seg), torch.empty(c, h, w, dtype=torch.int32, device=dev)))
torch.cuda.synchronize()
