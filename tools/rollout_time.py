import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
import helpers as H, bench as Bn
from jaxmarl_hft_b200 import env as E
mac = H.load_mac("2_player_fq_fqc"); ld = Bn._load_day(mac)
B, T = 16384, 64
env = E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=1)
p = env.default_params; obs, st = env.reset(None, p)
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.randint(0, env.action_spaces[t].n, (T, B, 1), generator=g, device="cuda", dtype=torch.int32) for t in range(2)]
for _ in range(2): env.rollout(st, acts, T, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): env.rollout(st, acts, T, p)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"rollout kernel: {ms:.2f} ms per {T}-step rollout of {B} envs = {B*T/ms*1e3:.3e} env-steps/s (incl. {T} PRNG draws)")
e0.record()
for _ in range(5): env.rollout(st, acts, T, p, draw=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"rollout kernel, draws reused: {ms:.2f} ms = {B*T/ms*1e3:.3e} env-steps/s")
