import numpy as np, sys
sys.path.insert(0, ".")
from oracle import orb as oorb
from visual_odometry_ros_b200 import capi, synth
ctx = capi.Context(device=0, max_w=1241, max_h=376, n_slots=2, max_feat=4096)
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (200, 300), dtype=np.uint8)
thr = 15
ctx.upload_image(0, img)
P, R, O = ctx.orb_detect(0, thr)
pyr = oorb.pyramid(img)
edge = 31
for lv in range(8):
    g = ctx.orb_read_level(lv, 0)
    print("level", lv, g.shape, pyr[lv].shape, "img equal", np.array_equal(g, pyr[lv]))
    h, w = pyr[lv].shape
    if w <= 2 * edge or h <= 2 * edge: continue
    so = oorb.fast_scores(pyr[lv], thr)
    sg = ctx.orb_read_level(lv, 1)
    band = (slice(edge - 1, h - edge + 1), slice(edge - 1, w - edge + 1))
    print("   score equal in band", np.array_equal(sg[band].astype(int), so[band]), int((sg[band].astype(int) != so[band]).sum()))
    xs, ys, sc = oorb.fast_detect(pyr[lv], thr)
    inb = (xs >= edge) & (xs < w - edge) & (ys >= edge) & (ys < h - edge)
    ng = ctx.orb_read_level(lv, 2)
    inner = (slice(edge, h - edge), slice(edge, w - edge))
    ref = np.zeros((h, w), int); ref[ys[inb], xs[inb]] = sc[inb]
    print("   nms equal", np.array_equal(ng[inner].astype(int), ref[inner]), int((ng[inner] > 0).sum()), int(inb.sum()))
Po, Ro, Oo = oorb.detect(img, thr)
print("counts gpu", np.bincount(O, minlength=8), "oracle", np.bincount(Oo, minlength=8), "quota", oorb.features_per_level())
key = lambda P, R, O: sorted((int(o), float(p[1]), float(p[0]), float(r)) for p, r, o in zip(P, R, O))
a, b = key(P, R, O), key(Po, Ro, Oo)
sa, sb = set(a), set(b)
print("only gpu", sorted(sa - sb)[:5], "only oracle", sorted(sb - sa)[:5])
so = oorb.fast_scores(img, thr)
sg = ctx.orb_read_level(0, 1).astype(int)
bad = np.argwhere(sg[30:170, 30:270] != so[30:170, 30:270])[:12] + 30
for y, x in bad:
    print(y, x, "gpu", sg[y, x], "oracle", so[y, x])
print("gpu nonzero", int((sg[30:170, 30:270] > 0).sum()), "oracle nonzero", int((so[30:170, 30:270] > 0).sum()))
print("mismatch where gpu==0:", int(((sg != so) & (sg == 0))[30:170, 30:270].sum()), " where oracle==0:", int(((sg != so) & (so == 0))[30:170, 30:270].sum()))
