"""How many outliers have an EMPTY vote region (no valid pixel in their cross region)?  They could be skipped with a prefix-sum
precount instead of being traversed in each of the five voting iterations."""
import sys
import numpy as np
sys.path.insert(0, '.')
import tea_stereo_matching_b200 as t
from tea_stereo_matching_b200 import _native as N
from tea_stereo_matching_b200.synth import synth_v1

def stats(left, right, D, name):
    r = t.StageRunner(left, right, D)
    r.run(N.STAGE_PREP | N.STAGE_INIT | N.STAGE_AGGREGATE | N.STAGE_SCANLINE | N.STAGE_LRC)
    arms = r.arms(0).astype(np.int32)  # up, down, left, right
    H, W = arms.shape[:2]
    for it in range(5):
        disp = r.disp()
        valid = (disp >= 0).astype(np.int32)
        V = np.zeros((H + 1, W), np.int32); V[1:] = np.cumsum(valid, axis=0)
        ys, xs = np.nonzero(disp < 0)
        cnt = np.zeros(len(ys), np.int64)
        if it % 2 == 1 or True:
            pass
        # vertical-first form (columns x-l..x+r, each with its own up/down): the k_vote_pass_a<false> geometry
        l, rr = arms[ys, xs, 2], arms[ys, xs, 3]
        for o in range(-34, 35):
            m = (o >= -l) & (o <= rr)
            c = np.clip(xs + o, 0, W - 1)
            up, dn = arms[ys, c, 0], arms[ys, c, 1]
            cnt += np.where(m, V[np.clip(ys + dn + 1, 0, H), c] - V[np.clip(ys - up, 0, H), c], 0)
        print(f"{name} it{it}: outliers {len(ys)} ({100 * len(ys) / (H * W):.1f} % of px), empty region {np.mean(cnt == 0) * 100:.1f} %, 1..20 votes {np.mean((cnt > 0) & (cnt <= 20)) * 100:.1f} %, > 20 votes {np.mean(cnt > 20) * 100:.1f} %", flush=True)
        r.run(N.STAGE_VOTE, it)
    r.close()

l, rg = synth_v1(1080, 1920, 192, seed=1000)
stats(l, rg, 192, "C3 synthetic")
z = np.load("tests/golden/pair_0600_320x180.npz")
stats(z["left"], z["right"], 48, "0600 320x180")
