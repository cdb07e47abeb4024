"""TEST INFRASTRUCTURE ONLY -- freezes golden vectors from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Writes tests/golden/*.npz.  The GPU box has no reference tree, so the -m gpu tests and the oracle
tests read only these files.  Every array is produced by executing reference code
(src/utils/mv_utils.py, src/utils/clip_utils.py, src/utils/pointcloud_utils.py,
src/vilgod/lidar_frame.py, src/vilgod/zero_shot_detector.py:389-415 glue, third_party/CLIP/clip/*)
through oracle/ref_harness.py; inputs come from vilgod_b200.synthetic with fixed seeds.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import time

import numpy as np
import torch

from oracle import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


def weights_fingerprint(sd):
    h = hashlib.sha256()
    sums = []
    for k in sorted(sd):
        a = sd[k].detach().float().contiguous().numpy()
        h.update(k.encode())
        h.update(a.tobytes())
        sums.append(float(a.astype(np.float64).sum()))
    return h.hexdigest(), np.asarray(sums)


def golden_tables(clipw):
    from src.utils import mv_utils
    out = {}
    for V in (4, 6, 10):
        out[f"rot{V}"] = rh.make_reference_projection(V).rot_mat.numpy()
    proj = rh.make_reference_projection(4)
    out["gauss"] = proj.grid2image.conv.weight.data.numpy().reshape(3, 3)
    out["conv_bias"] = proj.grid2image.conv.bias.data.numpy()
    out["text_features"] = clipw.text_features.detach().float().numpy()
    out["text_tokens0"] = clipw.text_tokenized[0].numpy()
    out["class_list"] = np.asarray(rh.CLASS_LIST)
    out["class_mapped"] = np.asarray([rh.CLASS_MAPPING[c] for c in rh.CLASS_LIST])
    _save("tables.npz", **out)


def golden_canon():
    from src.utils import pointcloud_utils as pu
    from vilgod_b200 import synthetic
    pts, off, _ = synthetic.make_clusters_raw(12, n_min=10, n_max=600, seed=11)
    a = 0.37
    T = np.eye(4)
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    T[:3, 3] = [1.5, -2.0, 0.3]
    ref = [pu.transform_cluster_points_to_origin(pu.apply_transform(pts[off[c]:off[c + 1]], T))
           for c in range(len(off) - 1)]
    _save("canon.npz", raw=pts, offsets=off, transform=T, canon_f64=np.concatenate(ref),
          canon_f32=np.concatenate(ref).astype(np.float32))


def _special_clusters():
    """Canonicalised clusters chosen to hit the rasteriser's corners: the bmm naive/BLAS switch
    (9N < 400), duplicated points, points exactly on the bounding cube, large N."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(3)
    clusters = []
    for n in (10, 17, 44, 45, 64, 100, 333, 1000, 2048, 4096):
        p, o = synthetic.make_clusters(1, n_min=n, n_max=n, rng=rng)
        clusters.append(p)
    p, _ = synthetic.make_clusters(1, n_min=200, n_max=200, rng=rng)
    clusters.append(np.concatenate([p, p[:50], p[:50]]))                 # duplicates
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, 8)] * 3), -1).reshape(-1, 3)
    clusters.append((g * [0.5, 1.0, 0.25] + [0, 0, 1]).astype(np.float32))  # lattice on cell edges
    off = np.zeros(len(clusters) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(c) for c in clusters])
    return np.concatenate(clusters).astype(np.float32), off


def golden_projection():
    from src.utils import mv_utils
    pts, off = _special_clusters()
    C = len(off) - 1
    out = dict(points=pts, offsets=off)
    for V in (4, 6, 10):
        proj = rh.make_reference_projection(V)
        rotated, cells, vals, dens, u8 = [], [], [], [], []
        for c in range(C):
            p = torch.from_numpy(pts[off[c]:off[c + 1]])[None]
            rp = proj.point_transform(torch.repeat_interleave(p, V, dim=0),
                                      proj.rot_mat.repeat(1, 1, 1))
            rotated.append(rp.numpy().copy())
            grid = mv_utils.points2grid(rp.clone(), proj.resolution, proj.depth, proj.obj_ratio,
                                        proj.depth_bias).squeeze()          # [V,D,X,Y]
            g = grid.permute(0, 1, 3, 2).contiguous().numpy()               # -> [V,D,Y,X]
            nz = np.flatnonzero(g)
            cells.append(nz.astype(np.int32))
            vals.append(g.reshape(-1)[nz])
        res = rh.reference_classification(proj, None,
                                          [pts[off[c]:off[c + 1]] for c in range(C)])
        # reference densified image is [x, y]-indexed (mv_utils.py:125); store as [y, x]
        out[f"dens{V}"] = np.ascontiguousarray(
            res["densified"].reshape(C, V, 110, 110).transpose(0, 1, 3, 2))
        out[f"u8_{V}"] = res["u8"].reshape(C, V, 224, 224)
        out[f"rotated{V}"] = np.concatenate([r.transpose(1, 0, 2).reshape(-1, V * 3)
                                             for r in rotated])            # [sum N, V*3]
        out[f"grid_cells{V}"] = np.concatenate(cells)
        out[f"grid_vals{V}"] = np.concatenate(vals)
        out[f"grid_counts{V}"] = np.asarray([len(c) for c in cells], dtype=np.int32)
    _save("projection.npz", **out)
    return out


def golden_projection224():
    """BASELINE.json configs[3] sweeps R in {112, 224}: the reference is parametric in `resolution`
    (src/utils/mv_utils.py:91-127), so the same loop at R = 224 gives 222x222 densified images."""
    from src.utils import mv_utils
    pts, off = _special_clusters()
    pick = [1, 5, 8, 11]                      # 17, 100, 2048 points and the cell-edge lattice
    clusters = [pts[off[c]:off[c + 1]] for c in pick]
    off2 = np.zeros(len(pick) + 1, dtype=np.int32)
    off2[1:] = np.cumsum([len(c) for c in clusters])
    V = 4
    proj = rh.make_reference_projection(V, resolution=224)
    cells, vals = [], []
    for p in clusters:
        t = torch.from_numpy(p)[None]
        rp = proj.point_transform(torch.repeat_interleave(t, V, dim=0), proj.rot_mat.repeat(1, 1, 1))
        grid = mv_utils.points2grid(rp.clone(), proj.resolution, proj.depth, proj.obj_ratio,
                                    proj.depth_bias).squeeze()
        g = grid.permute(0, 1, 3, 2).contiguous().numpy()
        nz = np.flatnonzero(g)
        cells.append(nz.astype(np.int32))
        vals.append(g.reshape(-1)[nz])
    res = rh.reference_classification(proj, None, clusters)
    dens = np.ascontiguousarray(res["densified"].reshape(len(pick), V, 222, 222).transpose(0, 1, 3, 2))
    _save("projection224.npz", points=np.concatenate(clusters), offsets=off2,
          grid_cells=np.concatenate(cells), grid_vals=np.concatenate(vals),
          grid_counts=np.asarray([len(c) for c in cells], dtype=np.int32),
          u8=res["u8"].reshape(len(pick), V, 224, 224), dens_c1=dens[1], dens_c2v0=dens[2, 0])


def golden_vit(clip_plain, clip_ln, u8_images):
    """8 depth images through the reference preprocess + encode_image + scoring, for the plain
    random-init checkpoint and for one whose LayerNorm weights/biases were perturbed."""
    from PIL import Image
    out = dict(u8=u8_images)
    pil = [Image.fromarray(np.repeat(a[..., None], 3, axis=2)) for a in u8_images]
    for tag, cw in (("plain", clip_plain), ("ln", clip_ln)):
        with torch.no_grad():
            x = torch.cat([cw.preprocess(p).unsqueeze(0) for p in pil])
            vis = cw.model.visual
            stages = {}
            hooks = [vis.ln_pre.register_forward_hook(lambda m, i, o: stages.__setitem__("ln_pre", o)),
                     vis.transformer.resblocks[0].register_forward_hook(
                         lambda m, i, o: stages.__setitem__("block0", o)),
                     vis.transformer.resblocks[11].register_forward_hook(
                         lambda m, i, o: stages.__setitem__("block11", o))]
            feats = cw.model.encode_image(x)
            for h in hooks:
                h.remove()
            f = feats / feats.norm(dim=-1, keepdim=True)
            logits = 100.0 * f @ cw.text_features.T
            names, scores = cw.predict_clip_labels(pil)
        out[f"{tag}_pre_x_sample"] = x[:, :, 100, 90:110].numpy()
        out[f"{tag}_ln_pre"] = stages["ln_pre"][:, :3, :].numpy()                  # [B,3 tok,768]
        out[f"{tag}_block0"] = stages["block0"].permute(1, 0, 2)[:, :3, :].numpy()   # LND -> NLD
        out[f"{tag}_block11"] = stages["block11"].permute(1, 0, 2)[:, :3, :].numpy()
        if tag == "ln":
            # every token of the first four images (fp16-compressed: 2^-11 relative, far below the
            # stated tolerances): a wrong row statistic or a ragged-tile bug at tokens 3..196 must
            # not hide behind the three rows above
            out["ln_block0_full"] = stages["block0"].permute(1, 0, 2)[:4].numpy().astype(np.float16)
            out["ln_block11_full"] = stages["block11"].permute(1, 0, 2)[:4].numpy().astype(np.float16)
        out[f"{tag}_feats"] = feats.numpy()
        out[f"{tag}_logits"] = logits.numpy()
        out[f"{tag}_probs"] = logits.softmax(dim=-1).numpy()
        out[f"{tag}_names"] = np.asarray(names)
        out[f"{tag}_scores"] = np.asarray(scores, dtype=np.float32)
        sha, sums = weights_fingerprint(cw.model.visual.state_dict())
        out[f"{tag}_weights_sha256"] = np.asarray(sha)
        out[f"{tag}_weights_sums"] = sums
    _save("vit.npz", **out)


class _Det:
    def __init__(self):
        for k in ("object_class", "object_class_score", "object_class_predictions",
                  "object_class_predictions_score", "object_class_predictions_detailed"):
            setattr(self, k, None)
        self.depth_image = None

    def add_object_entry(self, entry_name, key, data):
        if getattr(self, entry_name) is None:
            setattr(self, entry_name, {})
        getattr(self, entry_name)[key] = data


class _Frame:
    def __init__(self, n):
        self.detections = [_Det() for _ in range(n)]


def reference_vote(names, detailed, scores):
    from src.vilgod import lidar_frame
    fr = _Frame(len(names))
    lidar_frame.LidarFrame.update_object_classes(fr, names, detailed, scores, [True] * len(names),
                                                 key="k", aggregation="voting")
    return (np.asarray([d.object_class["k"] for d in fr.detections]),
            np.asarray([d.object_class_score["k"] for d in fr.detections], dtype=np.float32))


def golden_vote():
    rng = np.random.default_rng(5)
    mapped = np.asarray([rh.CLASS_MAPPING[c] for c in rh.CLASS_LIST])
    out = {}
    for V in (4, 6, 10):
        C = 300
        idx = rng.integers(0, 24, size=(C, V))
        idx[: C // 3] = rng.integers(0, 24, size=(C // 3, 1)) + rng.integers(0, 2, size=(C // 3, V))
        idx = np.clip(idx, 0, 23)
        scores = rng.uniform(0.04, 0.9, size=(C, V)).astype(np.float32)
        scores[::7] = 0.25                                        # exact score ties
        names = mapped[idx]
        detailed = np.asarray(rh.CLASS_LIST)[idx]
        vn, vs = reference_vote(names, detailed, scores)
        out[f"idx{V}"], out[f"scores{V}"] = idx.astype(np.int32), scores
        out[f"voted_name{V}"], out[f"voted_score{V}"] = vn, vs
    _save("vote.npz", **out)


def golden_e2e(clipw, num_clusters=128, V=6, name="e2e.npz", seed=None):
    """BASELINE.json configs[0]: one synthetic Waymo-shaped frame, 128 clusters <= 2048 points,
    6 views, fp32 CLIP on CPU, exactly the loop of zero_shot_detector.py:389-416.  With V = 10 and
    another seed: a slice of configs[1] (the benchmarked workload)."""
    from vilgod_b200 import synthetic
    pts, off = synthetic.make_clusters(num_clusters, seed=synthetic.DEFAULT_SEED if seed is None else seed)
    proj = rh.make_reference_projection(V)
    t = time.time()
    res = rh.reference_classification(proj, clipw, [pts[off[c]:off[c + 1]]
                                                    for c in range(num_clusters)])
    dt = time.time() - t
    print(f"  reference classification of {num_clusters} clusters x {V} views: {dt:.1f} s "
          f"({num_clusters / dt:.2f} clusters/s, {os.cpu_count()} cores)")
    names = res["names"].reshape(num_clusters, V)
    scores = res["scores"].reshape(num_clusters, V)
    mapped = np.vectorize(rh.CLASS_MAPPING.get)(names)
    vn, vs = reference_vote(mapped, names, scores)
    # per-image probabilities need a second pass (predict_clip_labels only returns top-1)
    from PIL import Image
    with torch.no_grad():
        probs = []
        for i in range(0, len(res["u8"]), 50):
            pil = [Image.fromarray(np.repeat(a[..., None], 3, axis=2)) for a in res["u8"][i:i + 50]]
            x = torch.cat([clipw.preprocess(p).unsqueeze(0) for p in pil])
            f = clipw.model.encode_image(x)
            f = f / f.norm(dim=-1, keepdim=True)
            probs.append(torch.cat([(100.0 * f @ clipw.text_features.T), f], dim=1))
        pf = torch.cat(probs).numpy()
    import zlib
    _save(name, points=pts, offsets=off,
          u8_crc32=np.asarray([zlib.crc32(a.tobytes()) for a in res["u8"]], dtype=np.uint32),
          u8_sum=res["u8"].reshape(len(res["u8"]), -1).sum(axis=1).astype(np.uint32),
          u8_first8=res["u8"][:8], logits=pf[:, :24], feats=pf[:, 24:].astype(np.float16),
          names=names, scores=scores, voted_name=vn, voted_score=vs,
          ref_seconds=np.asarray(dt), ref_cores=np.asarray(os.cpu_count()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    rh.install_shims()
    torch.set_num_threads(os.cpu_count())
    tmp = tempfile.mkdtemp(prefix="vilgod_ckpt_")
    plain_dir, ln_dir = os.path.join(tmp, "plain"), os.path.join(tmp, "ln")
    os.makedirs(plain_dir), os.makedirs(ln_dir)
    print("checkpoint ...")
    rh.make_random_checkpoint(os.path.join(plain_dir, "ViT-B-16.pt"), seed=1234)
    sd = torch.load(os.path.join(plain_dir, "ViT-B-16.pt"))
    from oracle import vit as ovit
    pert = ovit.perturb_layernorms({k[len("visual."):]: v for k, v in sd.items()
                                    if k.startswith("visual.")})
    for k, v in pert.items():
        sd["visual." + k] = v
    torch.save(sd, os.path.join(ln_dir, "ViT-B-16.pt"))
    clip_plain = rh.make_reference_clip(plain_dir)
    clip_ln = rh.make_reference_clip(ln_dir)
    only = [a[len("--only="):] for a in sys.argv if a.startswith("--only=")]
    want = lambda name: not only or name in only
    if want("tables"):
        print("tables ..."); golden_tables(clip_plain)
    if want("canon"):
        print("canon ..."); golden_canon()
    if want("projection") or want("vit"):
        print("projection ..."); pr = golden_projection()
    if want("vit"):
        print("vit ...")
        u8 = pr["u8_6"][[2, 5, 6, 7, 8, 9, 10, 11], [0, 1, 2, 3, 4, 5, 0, 1]]
        golden_vit(clip_plain, clip_ln, np.ascontiguousarray(u8))
    if want("vote"):
        print("vote ..."); golden_vote()
    if want("projection224"):
        print("projection224 ..."); golden_projection224()
    if want("e2e"):
        print("e2e ..."); golden_e2e(clip_plain)
    if want("e2e_cfg2"):
        print("e2e_cfg2 ..."); golden_e2e(clip_plain, num_clusters=96, V=10, name="e2e_cfg2.npz", seed=77)


if __name__ == "__main__":
    main()
