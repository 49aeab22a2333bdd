#!/usr/bin/env python
"""Development aid: find the first message at which the CUDA replay and the oracle disagree, print it and the rows.
    python tools/replay_debug.py [--books 40] [--msgs 500] [--events 30000]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers as H
from jaxmarl_hft_b200 import config as C, env as E
from oracle import lob_oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--books", type=int, default=40)
    ap.add_argument("--msgs", type=int, default=500)
    ap.add_argument("--events", type=int, default=30000)
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--bisect", action="store_true")
    a = ap.parse_args()
    oracle = lob_oracle.load()
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=a.events))
    bc = C.book_config(mac.world_config)
    be = E.BaseLOBEnv(mac.world_config, loaded=ld, device="cuda:0")
    P = be._params_np
    W = ld.starts.shape[0]
    B = a.books
    widx = np.arange(B) % W
    start = ld.starts[widx].astype(np.int64)

    def run(T):
        ra, rb, rt = P["init_asks"][widx].copy(), P["init_bids"][widx].copy(), P["init_trades"][widx].copy()
        ga, gb, gt = H.cuda_replay(bc, ra.copy(), rb.copy(), rt.copy(), ld.msgs, start, T)
        oracle.replay(bc, ra, rb, rt, ld.msgs, start, T)
        bad = [(ga[b] != ra[b]).any() or (gb[b] != rb[b]).any() or (gt[b] != rt[b]).any() for b in range(B)]
        return np.array(bad), (ra, rb, rt), (ga, gb, gt)

    if a.chunks > 1:   # the state persists across launches (BaseLOBEnv.step_env)
        import torch
        dev = torch.device("cuda:0")
        ra, rb, rt = P["init_asks"][widx].copy(), P["init_bids"][widx].copy(), P["init_trades"][widx].copy()
        ta, tb, tt = (torch.from_numpy(x.copy()).to(dev) for x in (ra, rb, rt))
        tm = torch.from_numpy(ld.msgs).to(dev)
        for k in range(a.chunks):
            off = start + k * a.msgs
            prev = (ra.copy(), rb.copy(), rt.copy())
            E.replay_books(bc, ta, tb, tt, tm, torch.from_numpy(off).to(dev), a.msgs)
            torch.cuda.synchronize()
            oracle.replay(bc, ra, rb, rt, ld.msgs, off, a.msgs)
            for name, r, g in zip(("asks", "bids", "trades"), (ra, rb, rt), (ta.cpu().numpy(), tb.cpu().numpy(), tt.cpu().numpy())):
                if (r != g).any() and a.bisect:
                    # bisect inside this chunk from the (equal) state before it
                    b = int(np.nonzero((r != g).any(axis=(1, 2)))[0][0])
                    lo_, hi_ = 0, a.msgs
                    def sub(T):
                        xa, xb, xt = (torch.from_numpy(x.copy()).to(dev) for x in prev)
                        E.replay_books(bc, xa, xb, xt, tm, torch.from_numpy(off).to(dev), T)
                        torch.cuda.synchronize()
                        ya, yb, yt = (x.copy() for x in prev)
                        oracle.replay(bc, ya, yb, yt, ld.msgs, off, T)
                        return (ya, yb, yt), (xa.cpu().numpy(), xb.cpu().numpy(), xt.cpu().numpy())
                    while hi_ - lo_ > 1:
                        mid = (lo_ + hi_) // 2
                        rr, gg = sub(mid)
                        if any((x[b] != y[b]).any() for x, y in zip(rr, gg)): hi_ = mid
                        else: lo_ = mid
                    rr, gg = sub(hi_)
                    print(f"chunk {k} book {b}: first divergence after message {hi_ - 1}: {ld.msgs[off[b] + hi_ - 1].tolist()}")
                    print("previous:", ld.msgs[off[b] + max(0, hi_ - 4): off[b] + hi_ - 1].tolist())
                    for nm, x, y in zip(("asks", "bids", "trades"), rr, gg):
                        for row in np.nonzero((x[b] != y[b]).any(axis=1))[0][:8]:
                            print(f"  {nm}[{row}] oracle {x[b][row].tolist()}  cuda {y[b][row].tolist()}")
                    pa = prev[0][b]; print("asks before chunk at that price:", pa[pa[:, 0] == ld.msgs[off[b] + hi_ - 1][3]].tolist())
                    pb = prev[1][b]; print("bids before chunk at that price:", pb[pb[:, 0] == ld.msgs[off[b] + hi_ - 1][3]].tolist())
                    return
                if (r != g).any():
                    bb, rows = np.nonzero((r != g).any(axis=2))
                    print(f"chunk {k}: {name} differ in {len(bb)} rows of books {sorted(set(bb.tolist()))[:10]}")
                    for b_, row in list(zip(bb, rows))[:10]:
                        print(f"  {name}[{b_}][{row}] oracle {r[b_][row].tolist()}  cuda {g[b_][row].tolist()}")
                    return
        print("no mismatch over", a.chunks, "chunks")
        return
    bad, ref, got = run(a.msgs)
    if not bad.any():
        print("no mismatch over", a.msgs, "messages")
        return
    b = int(np.argmax(bad))
    lo, hi = 0, a.msgs
    while hi - lo > 1:
        mid = (lo + hi) // 2
        bd, _, _ = run(mid)
        if bd[b]: hi = mid
        else: lo = mid
    bd, ref, got = run(hi)
    print(f"book {b}: first divergence after message index {hi - 1}: {ld.msgs[start[b] + hi - 1].tolist()}")
    print("previous messages:", ld.msgs[start[b] + max(0, hi - 4): start[b] + hi - 1].tolist())
    for name, r, g in zip(("asks", "bids", "trades"), ref, got):
        rows = np.nonzero((r[b] != g[b]).any(axis=1))[0]
        for row in rows[:8]:
            print(f"  {name}[{row}] oracle {r[b][row].tolist()}  cuda {g[b][row].tolist()}")


if __name__ == "__main__":
    main()
