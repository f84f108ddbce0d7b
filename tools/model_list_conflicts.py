"""CPU model of the shared-memory bank conflicts of the 7x7 lookup-list gather (roi_pool_gather_kernel) on a 64 x 64
map with three window tables: builds the lists of random 10..32 pixel RoIs exactly as roi_pool_desc_kernel does, groups
tasks into quarter-warps (8 consecutive (RoI, bin) indices share an LDS.128 wavefront when their 16-byte bank groups
differ) and prints the mean wavefronts per quarter-warp and slot -- without and with the (rejected) dealing of a bin's
positions over the slots.  profiles/r2_pool7_list_gather.md quotes its output (2.35 -> 1.93).    python tools/model_list_conflicts.py"""
import numpy as np, math
rng=np.random.default_rng(0)
H=W=64; WP=65; HWp=4160; smax=3
zero=smax*HWp; none=zero+1
def rnd(v): return int(math.floor(abs(v)+0.5)*(1 if v>=0 else -1))
def ranges(c1,c2,limit,P=7):
    s=rnd(c1); e=rnd(c2); b=np.float32(max(e-s+1,1))/np.float32(P)
    out=[]
    for i in range(P):
        lo=min(max(int(math.floor(np.float32(i)*b))+s,0),limit); hi=min(max(int(math.ceil(np.float32(i+1)*b))+s,0),limit)
        out.append((lo,hi))
    return out
def anchors(lo,hi,s):
    a=list(range(lo,hi-s+1,s))
    if (hi-lo)%s: a.append(hi-s)
    return a
def lists(box):
    x1,y1,x2,y2=box
    R=ranges(y1,y2,H); C=ranges(x1,x2,W)
    res=[]
    for ph in range(7):
        for pw in range(7):
            (y0,y1_),(x0,x1_)=R[ph],C[pw]; hh=y1_-y0; ww=x1_-x0
            o=[]
            if hh<=0 or ww<=0: o=[zero]
            else:
                s=min(smax,hh,ww)
                for ry in anchors(y0,y1_,s):
                    for cx in anchors(x0,x1_,s): o.append((s-1)*HWp+ry*WP+cx)
            o=(o+[none]*16)[:16]
            res.append(o)
    return res
def wf(col):  # wavefronts of one slot over 8 lanes: max multiplicity of distinct addresses per bank group
    groups={}
    for a in col:
        groups.setdefault(a&7,set()).add(a)
    return max(len(v) for v in groups.values())
tasks=[]
for _ in range(400):
    w=rng.uniform(10,32); h=rng.uniform(10,32); cx=rng.uniform(w/2,64-w/2); cy=rng.uniform(h/2,64-h/2)
    tasks+=lists((cx-w/2,cy-h/2,cx+w/2,cy+h/2))
tasks=np.array(tasks)
nq=len(tasks)//8
def total(t):
    tot=[0,0,0,0]; cnt=[0,0,0,0]
    for g in range(nq):
        q=t[g*8:(g+1)*8]
        for sl in range(4):
            if sl>=2 and all((x[2]>=zero and x[3]>=zero) for x in q): continue
            tot[sl]+=wf([int(x[sl]) for x in q]); cnt[sl]+=1
    return [tot[i]/max(cnt[i],1) for i in range(4)], cnt
print("no deal", total(tasks))
def deal(t):
    t=t.copy()
    for g in range(nq):
        for (p,q) in ((0,1),(2,3)):
            c0=[0]*8; c1=[0]*8
            for l in range(8):
                P=int(t[g*8+l][p]); Q=int(t[g*8+l][q]); vp=P<zero; vq=Q<zero
                keep=vp*c0[P&7]+vq*c1[Q&7]; swap=vp*c1[P&7]+vq*c0[Q&7]
                if swap<keep:
                    t[g*8+l][p],t[g*8+l][q]=Q,P
                    if vq: c0[Q&7]+=1
                    if vp: c1[P&7]+=1
                else:
                    if vp: c0[P&7]+=1
                    if vq: c1[Q&7]+=1
    return t
print("deal", total(deal(tasks)))
print("avg lookups", np.mean((tasks<zero).sum(1)))
