"""GPU parity of the fused projection kernel (through the C ABI) against the oracle and the
golden vectors frozen from the reference.  Bit-exact for masks / winners / uint8 vs the oracle,
1e-5 for densified images vs the reference."""
import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from oracle import projection as op

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from vilgod_b200 import _lib
    from vilgod_b200.engine import Engine
    cache = {}

    def get(V, mode=_lib.VG_ROTATE_TORCH_CPU, rot=None, tag=None, **kw):
        key = (V, mode, tag, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = Engine(num_views=V, rotate_mode=mode, rot_mat=rot, **kw)
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


def tiles_to_u8(tiles):
    B = tiles.shape[0]
    t = tiles.float().reshape(B, 14, 14, 16, 16).permute(0, 1, 3, 2, 4).reshape(B, 224, 224)
    return t.round().to(torch.uint8)


@pytest.mark.parametrize("V", [4, 6, 10])
def test_stage_isolated_scatter_is_bit_exact(golden, engines, V):
    """Same rotated fp32 points in (the reference's own bmm output), identity view: occupancy mask
    and fp32 winners must equal the reference's grid bit for bit."""
    g = golden["projection"]
    off = g["offsets"]
    rot = g[f"rotated{V}"].reshape(-1, V, 3)
    eng = engines(1, rot=torch.eye(3)[None], tag="identity")
    pts, offs = [], [0]
    for c in range(len(off) - 1):
        for v in range(V):
            pts.append(rot[off[c]:off[c + 1], v])
            offs.append(offs[-1] + off[c + 1] - off[c])
    out = eng.project(np.concatenate(pts), np.asarray(offs, np.int32), want_tiles=False,
                      want_grid=True)
    grid = out["grid"].cpu().numpy().reshape(len(off) - 1, -1)
    cells, vals, counts = g[f"grid_cells{V}"], g[f"grid_vals{V}"], g[f"grid_counts{V}"]
    pos = 0
    for c in range(len(off) - 1):
        nz = np.flatnonzero(grid[c])
        assert np.array_equal(nz, cells[pos:pos + counts[c]])
        assert np.array_equal(grid[c][nz], vals[pos:pos + counts[c]])
        pos += counts[c]


@pytest.mark.parametrize("V", [4, 6, 10])
def test_full_projection_against_golden_reference(golden, engines, V):
    """End to end with the torch-CPU rotation rule: grid bit-exact, densified <= 1e-5, uint8 within
    1 LSB on a tiny fraction of pixels (conv summation order is the only difference)."""
    g = golden["projection"]
    eng = engines(V)
    out = eng.project(g["points"], g["offsets"], want_u8=True, want_grid=True, want_densified=True)
    assert int(out["status"].abs().sum()) == 0
    C = len(g["offsets"]) - 1
    grid = out["grid"].cpu().numpy().reshape(C, -1)
    cells, vals, counts = g[f"grid_cells{V}"], g[f"grid_vals{V}"], g[f"grid_counts{V}"]
    pos = 0
    for c in range(C):
        nz = np.flatnonzero(grid[c])
        assert np.array_equal(nz, cells[pos:pos + counts[c]]), f"cluster {c}"
        assert np.array_equal(grid[c][nz], vals[pos:pos + counts[c]])
        pos += counts[c]
    dens = out["densified"].cpu().numpy().reshape(C, V, 110, 110)
    assert np.abs(dens - g[f"dens{V}"]).max() <= 1e-5
    u8 = out["u8"].cpu().numpy().reshape(C, V, 224, 224)
    diff = np.abs(u8.astype(int) - g[f"u8_{V}"].astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 3e-3
    assert torch.equal(tiles_to_u8(out["tiles"]), out["u8"])


@pytest.mark.parametrize("V", [6, 10])
def test_bit_exact_against_oracle_all_stages(engines, V):
    """Ragged synthetic clusters (N = 10 .. 16k): every stage equals the C oracle bit for bit."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(17)
    pts, off = synthetic.make_clusters(40, n_min=10, n_max=3000, rng=rng)
    big, boff = synthetic.make_clusters(2, n_min=16000, n_max=16384, rng=rng)
    pts = np.concatenate([pts, big])
    off = np.concatenate([off, boff[1:] + off[-1]]).astype(np.int32)
    eng = engines(V)
    out = eng.project(pts, off, want_u8=True, want_densified=True)
    dens_o, u8_o = opipe.project(pts, off, V, want_dens=True)
    assert np.array_equal(out["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
    assert np.array_equal(out["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)
    assert torch.equal(tiles_to_u8(out["tiles"]), out["u8"])


def test_large_clusters_reuse_the_spill_grids(engines):
    """Clusters above the shared-memory point cache (1408) keep their remaining quantised points in
    per-CTA global pools that are handed from CTA to CTA: more large images than pools, mixed with
    small ones, twice in a row, must still equal the oracle bit for bit."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(23)
    parts, offs = [], [np.zeros(1, np.int64)]
    for k in range(12):
        big, boff = synthetic.make_clusters(6, n_min=1409, n_max=9000, rng=rng)
        small, soff = synthetic.make_clusters(3, n_min=10, n_max=1408, rng=rng)
        for p_, o_ in ((big, boff), (small, soff)):
            parts.append(p_)
            offs.append(o_[1:].astype(np.int64) + offs[-1][-1])
    pts = np.concatenate(parts)
    off = np.concatenate(offs).astype(np.int32)
    eng = engines(6)                      # 72 large clusters x 6 views = 432 images > 296 pools
    a = eng.project(pts, off, want_u8=True, want_densified=True)
    b = eng.project(pts, off, want_u8=True)
    assert torch.equal(a["u8"], b["u8"]) and torch.equal(a["tiles"], b["tiles"])
    dens_o, u8_o = opipe.project(pts, off, 6, want_dens=True)
    assert np.array_equal(a["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
    assert np.array_equal(a["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)
    assert int(a["status"].abs().sum()) == 0


def test_rotation_modes(engines):
    from vilgod_b200 import _lib, synthetic
    pts, off = synthetic.make_clusters(12, n_min=10, n_max=80, seed=3)
    for mode, rule in ((_lib.VG_ROTATE_FUSED, "fused"), (_lib.VG_ROTATE_UNFUSED, "unfused")):
        eng = engines(6, mode=mode)
        out = eng.project(pts, off, want_tiles=False, want_u8=True)
        _, u8_o = opipe.project(pts, off, 6, rotate_rule=rule)
        assert np.array_equal(out["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)


def test_stamp_and_dense_paths_agree(engines):
    """Clusters up to 1024 points take the stamp path (max-pool applied at scatter time, Gaussian over
    the bounding box); asking for the raw grid forces the dense row-streaming path.  Both must give
    the same bits, and both equal the oracle."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(41)
    pts, off = synthetic.make_clusters(60, n_min=10, n_max=1024, rng=rng)
    edge, eoff = synthetic.make_clusters(3, n_min=1024, n_max=1024, rng=rng)     # the largest stamp size
    pts = np.concatenate([pts, edge])
    off = np.concatenate([off, eoff[1:] + off[-1]]).astype(np.int32)
    for V in (4, 10):
        eng = engines(V)
        a = eng.project(pts, off, want_u8=True, want_densified=True)                   # stamp
        b = eng.project(pts, off, want_u8=True, want_densified=True, want_grid=True)   # dense
        assert torch.equal(a["u8"], b["u8"]) and torch.equal(a["tiles"], b["tiles"])
        assert torch.equal(a["densified"], b["densified"])
        dens_o, u8_o = opipe.project(pts, off, V, want_dens=True)
        assert np.array_equal(a["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
        assert np.array_equal(a["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)


def test_fast_and_general_kernels_agree(monkeypatch):
    """The fast kernel (register-resident image, bounding-box buffer, walking emit; clusters up to 2048
    points) hands larger clusters over to the general kernel in list mode.  Every variant of it
    (VG_PROJ_VARIANT 1 = default, 2 = previous emit) must produce the bits of the general kernel alone (variant 0) and of the
    oracle, for tiles, uint8 (all views and first view only) and the densified tap."""
    from vilgod_b200 import synthetic
    from vilgod_b200.engine import Engine
    rng = np.random.default_rng(77)
    pts, off = synthetic.make_clusters(150, n_min=10, n_max=2048, rng=rng)
    edge, eoff = synthetic.make_clusters(4, n_min=2048, n_max=2048, rng=rng)      # the largest fast size
    over, ooff = synthetic.make_clusters(5, n_min=2049, n_max=5000, rng=rng)      # handed over
    tiny, toff = synthetic.make_clusters(6, n_min=1, n_max=3, rng=rng)
    parts, offs = [], [np.zeros(1, np.int64)]
    for p_, o_ in ((pts, off), (over, ooff), (edge, eoff), (tiny, toff)):
        parts.append(p_)
        offs.append(o_[1:].astype(np.int64) + offs[-1][-1])
    pts = np.concatenate(parts)
    off = np.concatenate(offs).astype(np.int32)
    V = 10
    outs = {}
    for variant in ("0", "1", "2"):
        monkeypatch.setenv("VG_PROJ_VARIANT", variant)
        eng = Engine(num_views=V)
        try:
            outs[variant] = eng.project(pts, off, want_u8=True, want_densified=True)
            again = eng.project(pts, off, want_u8=True)
            assert torch.equal(outs[variant]["tiles"], again["tiles"])
            only = eng.project(pts, off)           # the production call: tiles, no uint8 image
            assert torch.equal(outs[variant]["tiles"], only["tiles"])
        finally:
            eng.close()
    ref = outs["0"]
    st = ref["status"].cpu().numpy()
    ok = st == 0                       # single-point clusters have no extent: flagged, taps not written
    assert ok.sum() >= len(st) - 6
    okt = torch.as_tensor(ok, device=ref["status"].device)
    for variant in ("1", "2"):
        o = outs[variant]
        assert torch.equal(o["status"], ref["status"]), variant
        assert torch.equal(o["tiles"], ref["tiles"]), variant
        assert torch.equal(o["u8"], ref["u8"]), variant
        assert torch.equal(o["densified"].reshape(len(st), -1)[okt], ref["densified"].reshape(len(st), -1)[okt]), variant
    sel = np.flatnonzero(ok)
    sp = np.concatenate([pts[off[c]:off[c + 1]] for c in sel])
    so = np.zeros(len(sel) + 1, np.int32)
    so[1:] = np.cumsum([off[c + 1] - off[c] for c in sel])
    dens_o, u8_o = opipe.project(sp, so, V, want_dens=True)
    C = len(st)
    assert np.array_equal(outs["1"]["densified"].cpu().numpy().reshape(C, V, 110, 110)[ok], dens_o)
    assert np.array_equal(outs["1"]["u8"].cpu().numpy().reshape(C, V, 224, 224)[ok], u8_o)
    assert torch.equal(tiles_to_u8(outs["1"]["tiles"]), outs["1"]["u8"])
    # R = 224 (BASELINE configs[3]): the fast kernel's 512-thread instantiation against the general kernel
    r224 = {}
    for variant in ("0", "1", "2"):
        monkeypatch.setenv("VG_PROJ_VARIANT", variant)
        eng = Engine(num_views=6, resolution=224)
        try:
            r224[variant] = eng.project(pts, off, want_u8=True, want_densified=True)
            r224[variant]["tiles_only"] = eng.project(pts, off)["tiles"]     # the production call: no uint8 image
        finally:
            eng.close()
    for variant in ("1", "2"):
        a, b = r224[variant], r224["0"]
        assert torch.equal(a["status"], b["status"]), variant
        assert torch.equal(a["tiles"], b["tiles"]) and torch.equal(a["u8"], b["u8"]), variant
        assert torch.equal(a["tiles_only"], b["tiles"]), variant
        assert torch.equal(a["densified"].reshape(C, -1)[okt], b["densified"].reshape(C, -1)[okt]), variant


def test_r224_grid_against_reference_and_oracle(golden, engines):
    """BASELINE.json configs[3]: R = 224 grids (222x222 densified images).  Scatter winners bit-exact
    and densified <= 1e-5 against the reference's own run, every stage bit-exact against the oracle,
    on the stamp path (N <= 1024), the dense path, and a cluster beyond the shared-memory cache."""
    from vilgod_b200 import synthetic
    g = golden["projection224"]
    V = 4
    eng = engines(V, resolution=224)
    out = eng.project(g["points"], g["offsets"], want_u8=True, want_grid=True, want_densified=True)
    assert int(out["status"].abs().sum()) == 0
    C = len(g["offsets"]) - 1
    grid = out["grid"].cpu().numpy().reshape(C, -1)
    cells, vals, counts = g["grid_cells"], g["grid_vals"], g["grid_counts"]
    pos = 0
    for c in range(C):
        nz = np.flatnonzero(grid[c])
        assert np.array_equal(nz, cells[pos:pos + counts[c]]), f"cluster {c}"
        assert np.array_equal(grid[c][nz], vals[pos:pos + counts[c]])
        pos += counts[c]
    dens = out["densified"].cpu().numpy().reshape(C, V, 222, 222)
    assert np.abs(dens[1] - g["dens_c1"]).max() <= 1e-5
    assert np.abs(dens[2, 0] - g["dens_c2v0"]).max() <= 1e-5
    u8 = out["u8"].cpu().numpy().reshape(C, V, 224, 224)
    diff = np.abs(u8.astype(int) - g["u8"].astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 3e-3
    # oracle, all paths
    rng = np.random.default_rng(7)
    pts, off = synthetic.make_clusters(24, n_min=10, n_max=3000, rng=rng)
    big, boff = synthetic.make_clusters(1, n_min=9000, n_max=9000, rng=rng)
    pts = np.concatenate([pts, big])
    off = np.concatenate([off, boff[1:] + off[-1]]).astype(np.int32)
    a = eng.project(pts, off, want_u8=True, want_densified=True)
    n = np.diff(off)
    rot = op.view_rot_mats(V)
    for c in range(len(n)):
        p = pts[off[c]:off[c + 1]]
        d_o, u_o = op.project_batch(p, np.array([0, len(p)], np.int32), rot, R=224, fused=9 * len(p) >= 400)
        assert np.array_equal(a["densified"][c * V:(c + 1) * V].cpu().numpy(), d_o[0]), (c, len(p))
        assert np.array_equal(a["u8"][c * V:(c + 1) * V].cpu().numpy(), u_o[0]), (c, len(p))
    assert torch.equal(tiles_to_u8(a["tiles"]), a["u8"])


def test_cluster_beyond_the_point_pool(engines):
    """More points than the shared-memory cache plus the per-CTA pool hold (65,536): the tail is
    rotated and quantised again per depth slice.  Bit-exact against the oracle."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(5)
    big, boff = synthetic.make_clusters(1, n_min=70000, n_max=70000, rng=rng)
    small, soff = synthetic.make_clusters(2, n_min=50, n_max=500, rng=rng)
    pts = np.concatenate([big, small])
    off = np.concatenate([boff, soff[1:] + boff[-1]]).astype(np.int32)
    eng = engines(4)
    a = eng.project(pts, off, want_u8=True, want_densified=True)
    dens_o, u8_o = opipe.project(pts, off, 4, want_dens=True)
    assert np.array_equal(a["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
    assert np.array_equal(a["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)


def test_reciprocal_div_mode_matches_the_oracle_in_that_mode(engines):
    """VgConfig.div_mode = VG_DIV_RECIPROCAL evaluates `/ (1 + depth_bias)` the way torch-CUDA does
    (multiplication by the fp32 reciprocal).  No CUDA run of the reference exists to pin it, so the
    check is against the oracle restating the same rule; the default (true division) is what every
    golden vector pins."""
    from vilgod_b200 import _lib, synthetic
    pts, off = synthetic.make_clusters(30, n_min=10, n_max=4000, seed=12)
    eng = engines(6, div_mode=_lib.VG_DIV_RECIPROCAL)
    a = eng.project(pts, off, want_u8=True, want_densified=True)
    try:
        op.set_div_mode(1)
        dens_o, u8_o = opipe.project(pts, off, 6, want_dens=True)
    finally:
        op.set_div_mode(0)
    assert np.array_equal(a["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
    assert np.array_equal(a["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)
    b = engines(6).project(pts, off, want_densified=True, want_tiles=False)
    assert not torch.equal(a["densified"], b["densified"])      # the switch does change bits


def test_bf16_operand_build_emits_the_same_pixels(golden):
    """The bf16-operand library differs from the default fp16 one only in the tile element type."""
    from vilgod_b200.engine import Engine
    g = golden["projection"]
    e = Engine(num_views=6, operand_dtype="bf16")
    try:
        out = e.project(g["points"], g["offsets"], want_u8=True)
        assert out["tiles"].dtype == torch.bfloat16
        assert torch.equal(tiles_to_u8(out["tiles"]), out["u8"])
        diff = np.abs(out["u8"].cpu().numpy().reshape(g["u8_6"].shape).astype(int) - g["u8_6"].astype(int))
        assert diff.max() <= 1 and (diff != 0).mean() < 3e-3
    finally:
        e.close()


def test_degenerate_and_edge_clusters(engines):
    eng = engines(4)
    a = np.random.default_rng(0).normal(size=(50, 3)).astype(np.float32)
    same = np.ones((7, 3), np.float32)                       # zero extent
    one = np.array([[1.0, 2.0, 3.0]], np.float32)            # single point
    dup = np.repeat(a[:5], 10, axis=0)                       # heavy duplication, still has extent
    pts = np.concatenate([a, same, one, dup])
    off = np.array([0, 50, 57, 58, 58, 108], np.int32)       # includes an EMPTY cluster
    out = eng.project(pts, off, want_u8=True)
    st = out["status"].cpu().numpy()
    assert list(st) == [0, -4, -4, -4, 0]
    u8 = out["u8"].cpu().numpy().reshape(5, 4, 224, 224)
    assert (u8[1:4] == 0).all()
    _, u8_o = opipe.project(np.concatenate([a, dup]), np.array([0, 50, 100], np.int32), 4)
    assert np.array_equal(u8[[0, 4]], u8_o)
    # zero clusters is a no-op
    out0 = eng.project(np.zeros((0, 3), np.float32), np.zeros(1, np.int32))
    assert out0["tiles"].shape[0] == 0


def _adversarial_clusters():
    rng = np.random.default_rng(99)
    parts = []
    plane = rng.uniform(-1, 1, size=(64, 3)); plane[:, 2] = 0.25                 # no extent in z
    line = np.zeros((50, 3)); line[:, 0] = np.linspace(-2, 3, 50)                # extent in x only
    two = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0]])
    tiny = 5.0 + 1e-6 * rng.uniform(-1, 1, size=(40, 3))                         # fp32 grains
    far = 1e5 + rng.uniform(-0.5, 0.5, size=(200, 3))                            # large magnitude
    lattice = np.stack(np.meshgrid(*[np.arange(8.0)] * 3, indexing="ij"), -1).reshape(-1, 3)   # cell edges
    zcol = rng.normal(size=(300, 3)) * np.array([0.2, 0.2, 3.0])                 # z is the long axis
    box = rng.uniform(-1, 1, size=(2000, 3)) * np.array([2.3, 0.9, 0.8])
    shell = rng.normal(size=(5000, 3)); shell /= np.linalg.norm(shell, axis=1, keepdims=True)
    for a in (plane, line, two, tiny, far, lattice, zcol, box, shell):
        parts.append(np.asarray(a, np.float32))
    off = np.zeros(len(parts) + 1, np.int32)
    off[1:] = np.cumsum([len(a) for a in parts])
    return np.concatenate(parts), off


@pytest.mark.parametrize("depth_bias,obj_ratio", [(0.2, 0.8), (0.0, 0.8), (0.05, 1.0)])
def test_adversarial_geometry_and_slice_ties(depth_bias, obj_ratio):
    """Planar, linear, two-point, tiny, far-away, lattice-aligned and z-dominant clusters, under the
    reference's normalisation constants and two others.  depth_bias = 0 puts points into depth slice 0,
    whose value (1.0) ties with slice 1: the kernel's single-clear scheme has to notice and clear."""
    from vilgod_b200 import _lib
    from vilgod_b200.engine import Engine
    pts, off = _adversarial_clusters()
    V = 4
    eng = Engine(num_views=V, rotate_mode=_lib.VG_ROTATE_FUSED, depth_bias=depth_bias, obj_ratio=obj_ratio)
    try:
        out = eng.project(pts, off, want_u8=True, want_densified=True, want_grid=True)
        assert int(out["status"].abs().sum()) == 0
        dens_o, u8_o = op.project_batch_threaded(pts, off, op.view_rot_mats(V), fused=True, want_dens=True,
                                                 obj_ratio=obj_ratio, depth_bias=depth_bias)
        assert np.array_equal(out["densified"].cpu().numpy().reshape(dens_o.shape), dens_o)
        assert np.array_equal(out["u8"].cpu().numpy().reshape(u8_o.shape), u8_o)
        assert torch.equal(tiles_to_u8(out["tiles"]), out["u8"])
        # stage tap: the scatter-max grid of every (cluster, view) equals the oracle's points2grid
        grid = out["grid"].cpu().numpy().reshape(len(off) - 1, V, 8, 112, 112)
        rot = op.view_rot_mats(V)
        occupied = set()
        for c in (1, 5, 6):
            for v in range(V):
                q = op.rotate(pts[off[c]:off[c + 1]], rot[v], fused=True)
                g_o = op.points2grid(q, obj_ratio=obj_ratio, depth_bias=depth_bias)
                assert np.array_equal(grid[c, v], g_o.reshape(8, 112, 112)), (c, v)
                occupied |= set(np.flatnonzero(g_o.reshape(8, -1).max(axis=1) > 0).tolist())
        if depth_bias == 0.0:
            assert 0 in occupied          # the tie path was really exercised
    finally:
        eng.close()


def test_determinism_and_properties_at_scale(engines):
    """cfg2-sized slice (2 frames x ~300 clusters, 10 views): run-to-run bit stable (scatter-max is
    order independent); every image has background 255/254 only where untouched, and its minimum
    is exactly 0 (the max-depth pixel)."""
    from vilgod_b200 import synthetic
    frames = synthetic.make_frames(2, clusters_per_frame=300, seed=5)
    pts, off, _ = synthetic.concat_frames(frames)
    eng = engines(10)
    a = eng.project(pts, off, want_u8=True, want_densified=True)
    b = eng.project(pts, off, want_u8=True)
    assert torch.equal(a["tiles"], b["tiles"]) and torch.equal(a["u8"], b["u8"])
    dens = a["densified"]
    assert float(dens.min()) == 0.0 and float(dens.max()) == 1.0
    assert bool((dens.flatten(1).min(dim=1).values == 0).all())
    assert int(a["status"].abs().sum()) == 0
    # spot-check 16 random clusters against the oracle
    rng = np.random.default_rng(0)
    sel = rng.choice(len(off) - 1, 16, replace=False)
    sp = np.concatenate([pts[off[c]:off[c + 1]] for c in sel])
    so = np.zeros(17, np.int32)
    so[1:] = np.cumsum([off[c + 1] - off[c] for c in sel])
    _, u8_o = opipe.project(sp, so, 10)
    got = a["u8"].reshape(len(off) - 1, 10, 224, 224)[torch.as_tensor(sel)].cpu().numpy()
    assert np.array_equal(got, u8_o)


def test_reference_interface_get_img(golden, engines):
    """RealisticProjection.get_img mirror: [b,N,3] -> [b*V,3,110,110] in the reference orientation."""
    from vilgod_b200.reference_api import RealisticProjection
    g = golden["projection"]
    off = g["offsets"]
    proj = RealisticProjection(dict(resolution=112, depth=8, obj_ratio=0.8, depth_bias=0.2), num_views=4,
                               engine=engines(4))
    c = 8
    p = torch.from_numpy(g["points"][off[c]:off[c + 1]])[None].cuda()
    img = proj.get_img(p)
    assert img.shape == (4, 3, 110, 110)
    ref = np.transpose(g["dens4"][c], (0, 2, 1))
    assert np.abs(img[:, 0].cpu().numpy() - ref).max() <= 1e-5
    assert torch.equal(img[:, 0], img[:, 2])
    with pytest.raises(ValueError):
        proj.get_img(torch.ones(1, 5, 3).cuda())


def test_unsupported_configurations_are_rejected():
    from vilgod_b200.engine import Engine, VilgodError
    for kw in (dict(resolution=160), dict(pool_kernel=3), dict(pool_pad=2), dict(div_mode=7), dict(depth=6)):
        with pytest.raises(VilgodError):
            Engine(num_views=4, **kw)


def test_size_independent_properties_at_scale(engines):
    """Properties that hold bit for bit whatever the input (tests/test_oracle_properties.py pins them on the
    oracle), here on 2000 Waymo-shaped clusters x 10 views through the fast kernel and the hand-over list:
    point order and repeated points do not matter (scatter-max is order independent and idempotent), a
    power-of-two scale cancels exactly in (p - centre) / range, and a cluster's images do not depend on its
    neighbours in the packed batch."""
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(4242)
    pts, off = synthetic.make_clusters(2000, n_min=64, n_max=3000, rng=rng)     # >= 45 points: one bmm rule
    C = len(off) - 1
    eng = engines(10)
    base = eng.project(pts, off)
    assert int(base["status"].abs().sum()) == 0
    # permuted points + every point twice, cluster by cluster
    parts, noff = [], [0]
    for c in range(C):
        p = pts[off[c]:off[c + 1]]
        perm = rng.permutation(len(p))
        parts += [p[perm], p]
        noff.append(noff[-1] + 2 * len(p))
    dup = eng.project(np.concatenate(parts), np.asarray(noff, np.int32))
    assert torch.equal(dup["tiles"], base["tiles"])
    # power-of-two scales, a different one per cluster
    k = rng.integers(-6, 9, size=C)
    scale = np.repeat(np.exp2(k).astype(np.float32), np.diff(off))[:, None]
    scaled = eng.project((pts * scale).astype(np.float32), off)
    assert torch.equal(scaled["tiles"], base["tiles"])
    # clusters in reverse order: images move with their cluster
    order = np.arange(C)[::-1]
    rp = np.concatenate([pts[off[c]:off[c + 1]] for c in order])
    ro = np.zeros(C + 1, np.int32)
    ro[1:] = np.cumsum([off[c + 1] - off[c] for c in order])
    rev = eng.project(rp, ro)
    V = 10
    assert torch.equal(rev["tiles"].reshape(C, V, 196, 256).flip(0), base["tiles"].reshape(C, V, 196, 256))
