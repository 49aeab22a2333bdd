import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import jaxmarl_hft_b200
from jaxmarl_hft_b200 import config as C
from oracle import lob_oracle
import helpers as H
o=lob_oracle.load()
z=np.load('/root/repo/tests/golden/replay_adversarial.npz')
B,T,no,nt=int(z['B']),int(z['T']),int(z['no']),int(z['nt'])
bc=C.book_config(C.World_EnvironmentConfig(nOrders=no,nTrades=nt,type_4_interpretation=int(z['t4']),check_book_fill=bool(z['fill'])))
msgs=z['msgs']
# all prefixes at once: book (b,t) replays first t msgs of stream b
for b in range(B):
    ts=np.arange(1,T+1)
    nb=len(ts)
    found=False
    # run each prefix as a separate book with same start but different n? n_msgs is per launch -> loop launches in chunks: use bisect
    lo,hi=0,T
    def run(t):
        a=np.full((1,no,6),-1,np.int32); d=a.copy(); tr=np.full((1,nt,8),-1,np.int32)
        ra,rb,rt=a.copy(),d.copy(),tr.copy()
        st=np.array([b*T],np.int64)
        o.replay(bc,ra,rb,rt,msgs,st,t)
        ga,gb,gt=H.cuda_replay(bc,a,d,tr,msgs,st,t)
        return (np.array_equal(ga,ra) and np.array_equal(gb,rb) and np.array_equal(gt,rt)), (ra,rb,rt,ga,gb,gt)
    ok,_=run(T)
    if ok: print("stream",b,"ok"); continue
    while hi-lo>1:
        mid=(lo+hi)//2
        ok,_=run(mid)
        if ok: lo=mid
        else: hi=mid
    ok,(ra,rb,rt,ga,gb,gt)=run(hi)
    _,(pa,pb,pt,_,_,_)=run(lo)
    print("stream",b,"first bad prefix",hi,"msg",msgs[b*T+hi-1])
    print(" prev msgs", msgs[b*T+max(0,hi-4):b*T+hi-1].tolist())
    for name,r,g,p in (("asks",ra,ga,pa),("bids",rb,gb,pb),("trades",rt,gt,pt)):
        if not np.array_equal(r,g):
            idx=np.unique(np.argwhere(r!=g)[:,1])
            print(" ",name,"rows differ",idx.tolist())
            for i in idx[:4]: print("   row",i,"before",p[0,i].tolist(),"ref",r[0,i].tolist(),"got",g[0,i].tolist())
    print("  asks before (live):",[ (i,pa[0,i].tolist()) for i in range(no) if pa[0,i,0]!=-1 or pa[0,i,1]!=-1][:30])
    print("  bids before (live):",[ (i,pb[0,i].tolist()) for i in range(no) if pb[0,i,0]!=-1 or pb[0,i,1]!=-1][:30])
    break
