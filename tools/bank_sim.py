"""Numpy model of the shared-memory bank conflicts of K1's trilinear taps (CPU only, uses adell_aug_plan):\nwavefronts per LDS for the current lane mapping (32 lanes along axis 2, row pitch = box extent) and for a\n4 x 8 lane patch with other row pitches, on seg_all_affine-like items.  The current mapping reproduces ncu\n(2.15 modelled vs 2.18 measured); see DESIGN.md §9."""
import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, torch
from test_host_tile_policy import plan_items
from adell_mri_b200 import geometry
from adell_mri_b200.plan import BatchPlan

def fold(u, S):  # reflection about -0.5, S-0.5, then clamp
    y = (u + 0.5) / (2*S); f = y - np.rint(y); x = np.abs(f)*2*S - 0.5
    return np.clip(x, 0, S-1-1e-4)

def wavefronts(addr):
    # addr: [n_instr, 32] int
    bank = addr % 32
    out = np.zeros(addr.shape[0], int)
    for b in range(32):
        m = bank == b
        a = np.where(m, addr, -1)
        a.sort(axis=1)
        d = (np.diff(a, axis=1) != 0) & (a[:, 1:] >= 0)
        cnt = d.sum(1) + 0  # distinct non-negative addresses = transitions into new nonneg values
        # transitions count includes the -1 -> first value transition when there are -1s; handle:
        has_neg = (a[:, 0] < 0); has_pos = (a[:, -1] >= 0)
        n = np.where(has_pos, np.where(has_neg, cnt, cnt + 1), 0)
        out = np.maximum(out, n)
    return out

def sim(items, mode, pitch_rule, n=4000, seed=0, fold2=True):
    R = np.random.RandomState(seed)
    res = []
    for it in items:
        if int(it["kind"]) != 1: continue
        U0 = np.asarray(it["fp_U0"], float); D = np.asarray(it["fp_D"], float).reshape(3,3)
        sh = np.asarray(it["shear"]).reshape(2,16).astype(int)
        O = [int(x) for x in it["out_shape"]]; S = [int(x) for x in it["src_shape"]]
        box = [int(x) for x in it["tmap_box"]]
        pitch = pitch_rule(box[2]); b1 = box[1]
        if pitch_rule(0) == -1: b1 = (b1 + 3) & ~3
        p1 = pitch; p0 = b1 * pitch
        i = R.randint(0, O[0], n); j = R.randint(0, O[1] - 4, n)
        if mode == "row32":
            k = np.arange(32)[None, :] + np.zeros((n,1), int)
            jj = j[:, None] + 0*k
        else:
            G0 = R.randint(0, O[2] // 8, n)
            lane = np.arange(32)
            k = 8*G0[:, None] + (lane & 7)[None, :]
            jj = j[:, None] + (lane >> 3)[None, :]
        G = (k >> 3) & 15
        oi = i[:, None] - sh[0][G]; oj = jj - sh[1][G]
        u = [U0[a] + D[a,0]*oi + D[a,1]*oj + D[a,2]*k for a in range(3)]
        if fold2: u[2] = fold(u[2], S[2])
        c = [np.floor(x).astype(int) for x in u]
        addr = c[0]*p0 + c[1]*p1 + c[2]
        addr -= addr.min() - 0
        res.append(wavefronts(addr).mean())
    return np.array(res)

R = np.random.RandomState(1)
shape = (256, 256, 32)
mats = []
for b in range(16):
    ang = R.uniform(-1, 1, 3) * np.array([np.pi / 8, np.pi / 8, np.pi / 16])
    mats.append(geometry.compose_affine(rotate=ang[None])[0])
vols = [torch.zeros(shape) for _ in mats]
plan = BatchPlan(vols); plan.affine(np.stack(mats), "bilinear", "reflection")
it, tiles, info, _ = plan_items(plan, [shape]*len(mats))
print("boxes", [tuple(int(x) for x in b) for b in it["tmap_box"]][:6], [tuple(int(x) for x in b) for b in it["tile_dim"]][:6])
cur = sim(it, "row32", lambda b: b)
print("current (32 along k, pitch=box2): mean wavefronts/LDS", cur.mean().round(3), cur.round(2))
for name, rule in [("pitch 40", lambda b: 40 if b else 0), ("pitch 40, box1%4==0", lambda b: 40 if b else -1), ("pitch 36", lambda b: 36 if b else 0), ("pitch 32", lambda b: 32)]:
    new = sim(it, "patch", rule)
    print("4x8 patch,", name, new.mean().round(3), new.round(2))

# TMA SWIZZLE_128B on the current mapping (row pitch 32 floats = 128 B): the 16-byte chunk index of a cell is
# XORed with (row & 7); taps of lanes that share z but sit in different rows then fall into different banks.
def sim_swizzle(items, n=4000, seed=0):
    R = np.random.RandomState(seed)
    res = []
    for it in items:
        if int(it["kind"]) != 1: continue
        U0 = np.asarray(it["fp_U0"], float); D = np.asarray(it["fp_D"], float).reshape(3, 3)
        sh = np.asarray(it["shear"]).reshape(2, 16).astype(int)
        O = [int(x) for x in it["out_shape"]]; S = [int(x) for x in it["src_shape"]]
        box = [int(x) for x in it["tmap_box"]]
        if box[2] != 32: continue
        i = R.randint(0, O[0], n); j = R.randint(0, O[1], n)
        k = np.arange(32)[None, :] + np.zeros((n, 1), int)
        G = (k >> 3) & 15
        oi = i[:, None] - sh[0][G]; oj = j[:, None] - sh[1][G]
        u = [U0[a] + D[a, 0] * oi + D[a, 1] * oj + D[a, 2] * k for a in range(3)]
        u[2] = fold(u[2], S[2])
        c = [np.floor(x).astype(int) for x in u]
        row = c[0] * box[1] + c[1]
        row -= row.min()
        z = c[2]
        chunk = (z >> 2) ^ (row & 7)
        addr = row * 32 + chunk * 4 + (z & 3)
        res.append(wavefronts(addr).mean())
    return np.array(res)

sw = sim_swizzle(it)
print("32 along k, pitch 32, SWIZZLE_128B:", sw.mean().round(3), sw.round(2))
