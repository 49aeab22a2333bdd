import sys, dataclasses
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import jaxmarl_hft_b200
from jaxmarl_hft_b200 import config as C
from oracle import lob_oracle
import helpers as H
oracle=lob_oracle.load()
mac = H.load_mac("2_player_fq_fqc", nOrders=48, nTrades=20)
ld = H.load_for(mac, H.small_day(n_events=30000))
rng = np.random.default_rng(21)
bc = C.book_config(mac.world_config)
adv = H.adversarial_messages(rng, ld.msgs.shape[0], bc, tick=100)
adv[:, 3] = np.where(adv[:, 3] > 90_000, adv[:, 3] + 1_400_000, adv[:, 3])
keep = rng.random(ld.msgs.shape[0]) < 0.5
msgs = np.where(keep[:, None], ld.msgs, adv).astype(np.int32)
ld2 = dataclasses.replace(ld, msgs=np.ascontiguousarray(msgs))
B = 64
ref = H.OracleEnv(oracle, mac, ld2, B)
gpu = H.CudaEnv(mac, ld2, B, ref.params)
H.draw_prng(rng, ref.cfg, ref.arrays); gpu.set_inputs(ref.arrays)
ref.reset(); gpu.reset()
for s in range(66):
    H.draw_prng(rng, ref.cfg, ref.arrays); H.draw_actions(rng, ref.cfg, ref.arrays)
    gpu.set_inputs(ref.arrays)
    pre = {k: v.copy() for k, v in ref.arrays.items()}
    ref.step(n_threads=8); gpu.step()
    got = gpu.numpy()
    bad = [k for k in ref.arrays if not np.array_equal(ref.arrays[k], got[k], equal_nan=True) ]
    if bad:
        print("step", s, "mismatching leaves:", bad)
        k = "asks" if "asks" in bad else bad[0]
        envs = np.unique(np.argwhere(ref.arrays[k] != got[k])[:, 0])
        e = int(envs[0]); print(" envs", envs.tolist())
        print(" actions", [pre[f"actions{t}"][e].tolist() for t in range(2)], "perm", pre["perm"][e].tolist(), "step_counter", pre["step_counter"][e], "start", pre["start_index"][e])
        for name in ("asks", "bids"):
            r, g, p = ref.arrays[name][e], got[name][e], pre[name][e]
            rows = np.unique(np.argwhere(r != g)[:, 0])
            print(" ", name, "rows differ", rows.tolist())
            for i in rows[:6]: print("    row", i, "pre", p[i].tolist(), "ref", r[i].tolist(), "got", g[i].tolist())
            print("   pre live:", [(i, p[i].tolist()) for i in range(p.shape[0]) if (p[i] != -1).any()][:40])
        r, g = ref.arrays["best_asks"][e], got["best_asks"][e]
        d = np.argwhere((r != g).any(1))[:, 0]
        print("  best_asks first diff msg idx", d[:5].tolist(), "ref", r[d[:3]].tolist(), "got", g[d[:3]].tolist())
        r, g = ref.arrays["best_bids"][e], got["best_bids"][e]
        d2 = np.argwhere((r != g).any(1))[:, 0]
        print("  best_bids first diff msg idx", d2[:5].tolist(), "ref", r[d2[:3]].tolist(), "got", g[d2[:3]].tolist())
        off = int(pre["start_index"][e] + 100 * pre["step_counter"][e])
        first = int(min(list(d[:1]) + list(d2[:1]) + [111]))
        print("  data msgs around first diff (combined idx %d -> data idx %d):" % (first, first - 12))
        for j in range(max(0, first - 12 - 3), min(100, first - 12 + 2)): print("    ", j + 12, msgs[off + j].tolist())
        tr, tg = ref.arrays["trades"][e], got["trades"][e]
        print("  trades ref", tr[:6].tolist()); print("  trades got", tg[:6].tolist())
        break
else:
    print("all ok")
