"""Diagnostic (library built with EXTRA=-DRT_PROFILE_CLOCK): cycles every warp spends in the RT
kernel, as a map over the frame -- where the long-running blocks are.  python profiles/rt_block_cycles.py [workload]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as h, bench
b = importlib.import_module("computer-graphics_b200")
w = sys.argv[1] if len(sys.argv) > 1 else "rt_tess100k_4k"
kind, W, H, f = bench.WORKLOADS[w]
r = b.Renderer(0)
tris, sph = bench.scenes_rt(w)
cam = b.make_camera(bench.RT_CAM, f, h.identity_R(), W, H)
r.rt_upload_scene(tris, sph)
depth = torch.zeros((H, W), device="cuda"); index = torch.zeros((H, W), dtype=torch.int32, device="cuda"); rgb = torch.zeros((H, W, 3), device="cuda")
for _ in range(2):
    r.rt_render_device(cam, bench.RT_LIGHTS, 0, H, rgb.data_ptr(), depth.data_ptr(), index.data_ptr()); r.synchronize()
print("gpu_ms", r.stats()["gpu_ms"])
tot = depth.cpu().numpy(); ix = index.cpu().numpy()
recs = (ix & 0xfffff).astype(np.float32); cells = ((ix >> 20) & 2047).astype(np.float32)
prim = recs
# one value per warp (8x4 patch): take the max over the patch
wt = tot.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3)); wp = prim.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
bt = tot.reshape(H // 16, 16, W // 16, 16).max(axis=(1, 3))
us = lambda c: c / 1965.0
print("warp cycles (us): mean %.1f p50 %.1f p90 %.1f p99 %.1f p99.9 %.1f max %.1f" % tuple(us(x) for x in (wt.mean(), *np.percentile(wt, [50, 90, 99, 99.9, 100]))))
wc = cells.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
print("   shadow records streamed per warp: mean %.0f p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f" % (wp.mean(), *np.percentile(wp, [50, 90, 99, 99.9, 100])))
print("   light cells walked per warp: mean %.1f p90 %.0f p99 %.0f max %.0f" % (wc.mean(), *np.percentile(wc, [90, 99, 100])))
print("   correlation(cycles, records) = %.3f ; cycles per record: p50 %.1f" % (np.corrcoef(wt.ravel(), wp.ravel())[0, 1], np.median(wt[wp > 0] / wp[wp > 0])))
top = np.argsort(wt.ravel())[-8:][::-1]
for t in top:
    yy, xx = divmod(int(t), W // 8)
    patch = rgb[yy * 4: yy * 4 + 4, xx * 8: xx * 8 + 8].cpu().numpy().reshape(-1, 3)
    print(f"   slow warp at pixel ({xx * 8},{yy * 4}): {us(wt.ravel()[t]):.0f} us, {int(wp.ravel()[t])} records, {int(wc.ravel()[t])} cells; per lane: L1 tests {patch[:, 0].mean():.0f}, "
          f"candidates mean {patch[:, 1].mean():.0f} max {patch[:, 1].max():.0f}, exact evals mean {patch[:, 2].mean():.0f} max {patch[:, 2].max():.0f}")
g = rgb.cpu().numpy()
print("frame totals per pixel: L1 tests %.1f, candidates %.1f, shadow exact evals %.2f" % (g[..., 0].mean(), g[..., 1].mean(), g[..., 2].mean()))
print("block (slowest warp) us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(us(x) for x in (bt.mean(), *np.percentile(bt, [50, 90, 99, 100]))))
print("sum of warp time / (148 SMs x 16 warps) = %.3f ms" % (us(wt.sum()) / 1e3 / (148 * 16)))
for thr in (100, 200, 400, 800):
    print(f"blocks whose slowest warp takes > {thr} us: {(us(bt) > thr).sum()} of {bt.size}; their warp time share {wt[np.repeat(np.repeat(us(bt) > thr, 4, 0), 2, 1)].sum() / wt.sum():.2f}")
coarse = us(bt[: H // 16 // 9 * 9, : W // 16 // 16 * 16]).reshape(H // 16 // 9, 9, W // 16 // 16, 16).max(axis=(1, 3))
print("max block time (us) per 144x256-pixel region:")
for row in coarse:
    print("  ", " ".join(f"{v:6.0f}" for v in row))
