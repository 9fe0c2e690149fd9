"""Bank-conflict model of k_warp_fused's two shared-memory gathers (development evidence).

For warp instructions drawn from the bench's augmentation distribution it counts the data-pipe wavefronts of
  * the 16 pixel gathers (LDS.32 of a (B,G,R,mask) word; bank = (row * pitch + column) mod 32) for several footprint
    pitches and for 2-D lane shapes (32x1 = the kernel's, 16x2, 8x4, 4x8 destination pixels per instruction),
  * the weight-table gathers (LDS.128 per quarter-warp; 16-byte bank group = hash(ax, ay) & 7) for several hashes, and
    with a per-sample lane permutation (lane l takes pixel l*k mod 32).
It reproduces the measured conflict factors of the kernel (1.47 per pixel gather, 2.2 per weight gather).
    python tools/bank_sim.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmpe_b200
from oracle import gt_oracle as go
rng = np.random.RandomState(0)

def sample_rows(n_samples=40, rows_per_sample=200, shape='row'):
    """yield arrays (32,) of integer source X,Y (in 1/32 px fixed point) for the 32 lanes of one warp instruction"""
    out=[]
    for s in range(n_samples):
        smp = rmpe_b200.synth.gt_sample(s, 3, (368,368), True)
        flip,deg,crop,scale = smp['aug']
        M = go.affine_closed_form(flip,deg,crop,scale,smp['objpos'][0],smp['scale_provided'][0])
        A = np.vstack([M,[0,0,1]]); iA = np.linalg.inv(A)
        for r in range(rows_per_sample):
            tx = rng.randint(0,12); ty = rng.randint(0,12)
            x0 = tx*32; y0 = ty*32
            if shape=='row':
                ly = rng.randint(0,32)
                xs = x0 + np.arange(32); ys = np.full(32, y0+ly)
            elif shape=='8x4':
                bx = rng.randint(0,4); by = rng.randint(0,8)
                xs = x0 + bx*8 + (np.arange(32)%8); ys = y0 + by*4 + (np.arange(32)//8)
            elif shape=='16x2':
                bx = rng.randint(0,2); by = rng.randint(0,16)
                xs = x0 + bx*16 + (np.arange(32)%16); ys = y0 + by*2 + (np.arange(32)//16)
            elif shape=='4x8':
                bx = rng.randint(0,8); by = rng.randint(0,4)
                xs = x0 + bx*4 + (np.arange(32)%4); ys = y0 + by*8 + (np.arange(32)//4)
            sx = iA[0,0]*xs + iA[0,1]*ys + iA[0,2]
            sy = iA[1,0]*xs + iA[1,1]*ys + iA[1,2]
            X = np.floor(sx*32+0.5).astype(np.int64); Y = np.floor(sy*32+0.5).astype(np.int64)
            # skip rows entirely outside the source
            if (X>>5).max() < -2 or (X>>5).min() > 370 or (Y>>5).max() < -2 or (Y>>5).min() > 370: continue
            out.append((X,Y))
    return out

def wavefronts(banks, addrs, nb=32):
    """max over banks of #distinct addresses"""
    m=0
    for b in np.unique(banks):
        m=max(m, len(np.unique(addrs[banks==b])))
    return m

def pixel_conf(rows, pitch_mod, fn=None):
    tot=0;n=0
    for X,Y in rows:
        cx = (X>>5); cy=(Y>>5)
        cx = cx - cx.min(); cy = cy - cy.min()
        pitch = ((cx.max()+4+31)//32)*32 + pitch_mod
        for ky in range(4):
            for kx in range(4):
                addr = (cy+ky)*pitch + cx+kx
                tot += wavefronts(addr%32, addr); n+=1
    return tot/n

def weight_conf(rows, hashfn):
    tot=0;n=0
    for X,Y in rows:
        ax = X&31; ay = Y&31
        slot = ax*32+ay
        grp = hashfn(ax,ay)
        for q in range(4):
            sl = slice(8*q,8*q+8)
            tot += wavefronts(grp[sl], slot[sl]); n+=1
    return tot/n


def weight_conf_perm(rows_by_sample, ks=(1,3,5,7,9,11,13,15), hashfn=lambda ax,ay: ay&7):
    """per sample: choose the lane permutation l -> pixel (l*k mod 32) with the fewest weight wavefronts"""
    res=[]; base=[]
    for rows in rows_by_sample:
        best=None
        for k in ks:
            perm = (np.arange(32)*k)%32
            tot=0;n=0
            for X,Y in rows:
                ax=(X&31)[perm]; ay=(Y&31)[perm]
                slot=ax*32+ay; grp=hashfn(ax,ay)
                for q in range(4):
                    sl=slice(8*q,8*q+8); tot+=wavefronts(grp[sl],slot[sl]); n+=1
            v=tot/n
            if k==1: base.append(v)
            if best is None or v<best: best=v
        res.append(best)
    return np.mean(base), np.mean(res)

def sample_rows_by_sample(n_samples=60, rows_per_sample=60):
    out=[]
    for s in range(n_samples):
        smp = rmpe_b200.synth.gt_sample(s, 3, (368,368), True)
        flip,deg,crop,scale = smp['aug']
        M = go.affine_closed_form(flip,deg,crop,scale,smp['objpos'][0],smp['scale_provided'][0])
        A = np.vstack([M,[0,0,1]]); iA = np.linalg.inv(A)
        rows=[]
        for r in range(rows_per_sample):
            x0 = rng.randint(0,12)*32; y0 = rng.randint(0,12)*32+rng.randint(0,32)
            xs = x0+np.arange(32); ys=np.full(32,y0)
            sx = iA[0,0]*xs + iA[0,1]*ys + iA[0,2]; sy = iA[1,0]*xs + iA[1,1]*ys + iA[1,2]
            X = np.floor(sx*32+0.5).astype(np.int64); Y = np.floor(sy*32+0.5).astype(np.int64)
            if (X>>5).max() < -2 or (X>>5).min() > 370 or (Y>>5).max() < -2 or (Y>>5).min() > 370: continue
            rows.append((X,Y))
        if rows: out.append(rows)
    return out


if __name__=='__main__':
    for shape in ('row','16x2','8x4','4x8'):
        rows = sample_rows(shape=shape)
        print(shape, len(rows), 'pixel pitch0: %.3f'%pixel_conf(rows,0), ' '.join('p%d: %.3f'%(pm,pixel_conf(rows,pm)) for pm in (1,4,8,16)))
        print('   weights transposed(ay&7): %.3f  natural(ax&7): %.3f  sum: %.3f  xorfold: %.3f'%(
            weight_conf(rows, lambda ax,ay: ay&7), weight_conf(rows, lambda ax,ay: ax&7),
            weight_conf(rows, lambda ax,ay: (ax+ay)&7), weight_conf(rows, lambda ax,ay: (ax^ay^(ax>>3)^(ay>>3))&7)))
    rbs = sample_rows_by_sample()
    print('lane permutation, hash ay&7      base %.3f best-k %.3f' % weight_conf_perm(rbs))
    print('lane permutation, hash (ax+ay)&7 base %.3f best-k %.3f' % weight_conf_perm(rbs, hashfn=lambda ax, ay: (ax + ay) & 7))
