"""GPU parity of the ViT kernels (tcgen05 GEMM, fused attention, LayerNorm, head, vote) through the
C ABI.  Kernel-level checks use a plain torch fp32 reference of the same op on operand-rounded
inputs (every GEMM instantiation the tower launches has one); tower-level checks use the oracle (oracle/vit.py) and the golden vectors from the
reference.  Tolerances are stated next to each assert."""
import numpy as np
import pytest
import torch

from oracle import vit as ovit
from oracle import vote as ovote

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["f16", "bf16"])
def eng(request):
    from vilgod_b200.engine import Engine
    e = Engine(num_views=6, operand_dtype=request.param)
    yield e
    e.close()


def _ulp(e):
    """output rounding of one operand-typed value: bf16 keeps 8 significant bits, fp16 11"""
    return 2.0 ** -8 if e.operand_dtype == "bf16" else 2.0 ** -11


def _loaded_engine(golden, tag, operand_dtype="f16"):
    from vilgod_b200 import weights
    from vilgod_b200.engine import Engine
    e = Engine(num_views=6, operand_dtype=operand_dtype)
    sd = weights.random_init_visual_state_dict(1234)
    if tag == "ln":
        sd = weights.perturb_layernorms(sd)
    e.load_vit_weights(sd)
    e.set_text_features(golden["tables"]["text_features"])
    return e


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 768, 256), (197 * 3, 2304, 768),
                                   (1000, 3072, 768), (197 * 5 + 3, 768, 3072), (1, 256, 64),
                                   (148 * 128 * 2 + 77, 768, 768)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_tcgen05_gemm_against_torch(eng, M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(eng.op_torch_dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(eng.op_torch_dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    if epi == 0:
        out = eng.test_gemm(a, w, bias, 0).float()
        tol = _ulp(eng) * ref.abs().max().item() + 1e-3        # one rounding of the output
    elif epi == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
        out = eng.test_gemm(a, w, bias, 1).float()
        tol = _ulp(eng) * ref.abs().max().item() + 2e-3        # + tanh.approx in QuickGELU
    else:
        x0 = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + x0
        out = eng.test_gemm(a, w, bias, 2, out=x0.clone())
        tol = 2e-4 * max(1.0, K / 768)                          # fp32 accumulation order only
    err = (out - ref).abs().max().item()
    assert err <= tol, (err, tol)


RAGGED_M = [197 * 5 + 3, 148 * 256 + 77]


def _row_stats(x):
    """[M,3,2] per-row (sum, sum of squares) slots as the residual epilogues leave them: one slot per
    256-column tile of the 768-wide row."""
    t = x.float().view(x.shape[0], 3, 256)
    return torch.stack([t.sum(-1), (t * t).sum(-1)], dim=-1).contiguous()


@pytest.mark.parametrize("M", RAGGED_M)
@pytest.mark.parametrize("epi,N", [(0, 2304), (1, 3072)])
def test_layernorm_folded_gemm_variants_against_torch(eng, M, epi, N):
    """The instantiations the tower really runs for QKV (<bias, LNF>) and c_fc (<QuickGELU, LNF>):
    A is the RAW residual in the operand type, the weights carry the LayerNorm gain, and the epilogue
    applies rstd_i (acc - mu_i colsum_n) + c_n (model.py:157-163,190-191) -- against LayerNorm followed
    by the linear layer in fp32 torch, on ragged M and a residual with a DC offset."""
    K = 768
    g = torch.Generator(device="cuda").manual_seed(M + epi)
    x = torch.randn(M, K, device="cuda", generator=g) * 1.7 + 0.6
    xb = x.to(eng.op_torch_dtype)
    gamma = 1.0 + 0.2 * torch.randn(K, device="cuda", generator=g)
    beta = 0.2 * torch.randn(K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    wf = (w * gamma).to(eng.op_torch_dtype)                       # what fold_ln_kernel stores
    colsum = wf.float().sum(dim=1)
    cvec = w @ beta + b
    stats = _row_stats(xb)                                        # statistics of the values the GEMM sees
    out = eng.test_gemm_lnf(xb, wf, cvec, epi, stats, colsum=colsum).float()
    xf = xb.float()
    mu = xf.mean(dim=1, keepdim=True)
    var = (xf * xf).mean(dim=1, keepdim=True) - mu * mu
    ref = ((xf - mu) * torch.rsqrt(var + 1e-5)) @ wf.float().T + cvec
    if epi == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    tol = _ulp(eng) * ref.abs().max().item() + 3e-3
    assert (out - ref).abs().max().item() <= tol
    # and against the un-folded formulation (LayerNorm -> rounded operand -> linear): the fold may only
    # differ by operand rounding
    y = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5)
    ref2 = y @ w.T + b
    if epi == 1:
        ref2 = ref2 * torch.sigmoid(1.702 * ref2)
    assert (out - ref2).abs().max().item() <= (0.04 if eng.operand_dtype == "bf16" else 0.008) * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("M", RAGGED_M)
@pytest.mark.parametrize("K", [768, 3072])
def test_residual_gemm_variants_emit_copy_and_statistics(eng, M, K):
    """<resid, LNF, kWide> (out-proj, K = 768) and <resid, LNF, kDeep> (c_proj, K = 3072).  The tower
    keeps the residual stream as two operand-typed planes x = hi + lo that these epilogues update in
    place (the hook splits / merges fp32 around the production kernel): the updated residual, the hi
    plane (= the next GEMM's A operand) and the per-row (sum, sum of squares) slots, on ragged M."""
    N = 768
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(eng.op_torch_dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(eng.op_torch_dtype)
    b = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g) + 0.3
    if M < 2000:         # massive-activation channels, as real CLIP residual streams carry
        x0[:, 17] += 250.0
        x0[:, 403] -= 120.0
    stats = torch.full((M, 3, 2), float("nan"), device="cuda")
    x, xb = eng.test_gemm_lnf(a, w, b, 2, stats, x_inout=x0.clone())
    # the planes hold x0 to 2^-22 (fp16) / 2^-16 (bf16) relative before the update and x after it
    split = 2.0 ** -21 if eng.operand_dtype == "f16" else 2.0 ** -15
    ref = x0 + (a.float() @ w.float().T + b)
    assert (x - ref).abs().max().item() <= 2e-4 * max(1.0, K / 768) + split * ref.abs().max().item()
    # the hi plane is the rounded new residual (up to ties of the lo rounding)
    assert (xb != x.to(eng.op_torch_dtype)).float().mean().item() < 1e-3
    assert (xb.float() - x).abs().max().item() <= _ulp(eng) * x.abs().max().item()
    rs = _row_stats(x)
    assert torch.allclose(stats, rs, rtol=3e-5, atol=2e-3 + 2e-6 * rs.abs().max().item())   # fp32 sums of 256 terms, other order


@pytest.mark.parametrize("B", [1, 5, 300])
def test_patch_embedding_gemm_against_torch(eng, B):
    """<resid, wide, PATCH>: per-image 3-D TMA tiles, the [197,768] bias/position table as the added
    rows, token rows 1..196 written and the class-token row left alone."""
    g = torch.Generator(device="cuda").manual_seed(B)
    tiles = torch.randint(0, 256, (B, 196, 256), device="cuda", generator=g).to(eng.op_torch_dtype)
    w = (torch.randn(768, 256, device="cuda", generator=g) * 1e-3).to(eng.op_torch_dtype)
    table = torch.randn(197, 768, device="cuda", generator=g)
    x = torch.full((B, 197, 768), 7.0, device="cuda")
    eng.test_gemm_patch(tiles, w, table, x)
    ref = tiles.float() @ w.float().T + table[1:]
    assert (x[:, 1:] - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    assert bool((x[:, 0] == 7.0).all())


def test_layernorm_fold_survives_outlier_channels_and_dc_offset(golden):
    """Real CLIP residual streams carry a few massive-activation channels and non-zero row means;
    random-init weights do not.  Inject both after ln_pre-sized activations and compare the folded
    QKV GEMM with the unfused LayerNorm -> GEMM path kernel by kernel: the mu * colsum cancellation
    must stay inside the operand rounding of the unfused path."""
    from vilgod_b200.engine import Engine
    e = Engine(num_views=4)          # the default (fp16-operand) build
    try:
        M, K, N = 197 * 6, 768, 2304
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(M, K, device="cuda", generator=g) + 2.5          # DC offset
        x[:, 17] += 60.0                                                  # massive channels
        x[:, 403] -= 35.0
        gamma = 1.0 + 0.2 * torch.randn(K, device="cuda", generator=g)
        beta = 0.2 * torch.randn(K, device="cuda", generator=g)
        w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).half().float()
        b = torch.randn(N, device="cuda", generator=g) * 0.1
        xb = x.to(e.op_torch_dtype)
        wf = (w * gamma).to(e.op_torch_dtype)
        out = e.test_gemm_lnf(xb, wf, w @ beta + b, 0, _row_stats(xb), colsum=wf.float().sum(dim=1)).float()
        y = e.test_layernorm(x, gamma, beta)                                # unfused: LN kernel ...
        unf = e.test_gemm(y, w.to(e.op_torch_dtype), b, 0).float()          # ... then the plain GEMM
        ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ w.T + b
        err_fold = (out - ref).abs().max().item()
        err_unf = (unf - ref).abs().max().item()
        print(f"outlier residual: folded error {err_fold:.4f}, unfused error {err_unf:.4f}, |ref| max {ref.abs().max():.2f}")
        assert err_fold <= max(2.5 * err_unf, 0.02)
    finally:
        e.close()


@pytest.mark.parametrize("B", [1, 3, 16])
def test_fused_attention_against_torch(eng, B):
    g = torch.Generator(device="cuda").manual_seed(B)
    qkv = (torch.randn(B, 197, 2304, device="cuda", generator=g)).to(eng.op_torch_dtype)
    qkv[:, :, :768] *= 0.35    # keep logits in a realistic range (q arrives pre-scaled by 1/8)
    out = eng.test_attention(qkv).float()
    q, k, v = qkv.float().split(768, dim=-1)
    q = q.view(B, 197, 12, 64).transpose(1, 2)
    k = k.view(B, 197, 12, 64).transpose(1, 2)
    v = v.view(B, 197, 12, 64).transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(B, 197, 768)
    # P is rounded to the operand type before P.V and so is the output: one ulp relative each
    scale = 1.0 if eng.operand_dtype == "bf16" else 0.25
    assert (out - ref).abs().max().item() <= 2e-2 * scale
    assert (out - ref).abs().mean().item() <= 2e-3 * scale


def test_layernorm_against_torch(eng):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(197 * 4 + 5, 768, device="cuda", generator=g) * 3 + 0.7
    w = torch.randn(768, device="cuda", generator=g)
    b = torch.randn(768, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-5)
    out = eng.test_layernorm(x, w, b).float()
    assert (out - ref).abs().max().item() <= _ulp(eng) * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("operand_dtype", ["bf16", "f16"])
@pytest.mark.parametrize("tag", ["plain", "ln"])
def test_tower_stages_against_reference_golden(golden, tag, operand_dtype):
    """8 golden depth images: residual stream after ln_pre / block 0 / block 11 (rows 0..2) and the
    final embedding against the reference's fp32 run.  bf16 GEMM operands, fp32 everything else."""
    from vilgod_b200.engine import u8_to_tiles
    g = golden["vit"]
    e = _loaded_engine(golden, tag, operand_dtype)
    k = 1.0 if operand_dtype == "bf16" else 0.25      # fp16 operands: weights exact, 3 more bits
    try:
        tiles = u8_to_tiles(torch.from_numpy(g["u8"]).cuda(), e.op_torch_dtype)
        x = e.encode_score(tiles, stop_after_layer=-2)["x"][:, :3].cpu().numpy()
        # patch-embed with bf16 weights: K=256 products of integers <=255 with 2^-9 relative weight
        # error, then LayerNorm (unit variance output)
        assert np.abs(x - g[f"{tag}_ln_pre"]).max() <= 3e-2 * k
        x = e.encode_score(tiles, stop_after_layer=0)["x"][:, :3].cpu().numpy()
        assert np.abs(x - g[f"{tag}_block0"]).max() <= 6e-2 * k
        xf = e.encode_score(tiles, stop_after_layer=11)["x"].cpu().numpy()
        x = xf[:, :3]
        ref = g[f"{tag}_block11"]
        assert np.abs(x - ref).max() <= (0.03 * np.abs(ref).max() + 0.1) * k
        if tag == "ln":      # all 197 tokens of four images after the first and the last block
            x0 = e.encode_score(tiles, stop_after_layer=0)["x"][:4].cpu().numpy()
            r0 = g["ln_block0_full"].astype(np.float32)
            assert np.abs(x0 - r0).max() <= 6e-2 * k + 2.0 ** -11 * np.abs(r0).max()
            r11 = g["ln_block11_full"].astype(np.float32)
            assert np.abs(xf[:4] - r11).max() <= (0.03 * np.abs(r11).max() + 0.1) * k
        res = e.encode_score(tiles, want_logits=True)
        f_ref = g[f"{tag}_feats"] / np.linalg.norm(g[f"{tag}_feats"], axis=1, keepdims=True)
        cos = (res["feats"].cpu().numpy() * f_ref).sum(axis=1)
        assert cos.min() >= (0.9995 if operand_dtype == "bf16" else 0.99995), cos.min()
        # stated bf16 tolerance on logits (100 * cos units): 0.15 raw, 0.06 after removing the
        # per-image offset that soft-max ignores
        d = res["logits"].cpu().numpy() - g[f"{tag}_logits"]
        print(f"{operand_dtype}/{tag}: logit error raw {np.abs(d).max():.4f}, centred "
              f"{np.abs(d - d.mean(axis=1, keepdims=True)).max():.4f}, min feature cosine {cos.min():.6f}")
        assert np.abs(d).max() <= 0.15 * k, np.abs(d).max()
        assert np.abs(d - d.mean(axis=1, keepdims=True)).max() <= 0.06 * k
        assert np.abs(res["probs"].cpu().numpy() - g[f"{tag}_probs"]).max() <= 0.01
    finally:
        e.close()


@pytest.mark.parametrize("switches", [("VG_LN_UNFUSED",), ("VG_GEMM_NARROW",),
                                      ("VG_LN_UNFUSED", "VG_GEMM_NARROW")])
def test_alternate_kernel_paths_stay_correct(golden, monkeypatch, switches):
    """The A/B switches (read once in vg_create; "0" and "" mean off) select the unfused LayerNorm
    kernels and the 4-warp residual epilogues: each must meet the same stated tolerances against the
    reference as the production path."""
    from vilgod_b200.engine import u8_to_tiles
    g = golden["vit"]
    for name in switches:
        monkeypatch.setenv(name, "1")
    monkeypatch.setenv("VG_ATTN_TRACE", "0")        # "0" must read as off
    e = _loaded_engine(golden, "ln", "bf16")
    try:
        tiles = u8_to_tiles(torch.from_numpy(g["u8"]).cuda(), e.op_torch_dtype)
        res = e.encode_score(tiles, want_logits=True)
        d = res["logits"].cpu().numpy() - g["ln_logits"]
        assert np.abs(d).max() <= 0.15 and np.abs(d - d.mean(axis=1, keepdims=True)).max() <= 0.06
        assert np.abs(res["probs"].cpu().numpy() - g["ln_probs"]).max() <= 0.01
    finally:
        e.close()


def test_head_matches_oracle_given_same_residual(golden):
    """Isolate the fused head: feed the kernel's own block-11 residual stream to the oracle's
    ln_post / proj / normalise / score and compare at fp32 tolerance."""
    from vilgod_b200.engine import u8_to_tiles
    g = golden["vit"]
    e = _loaded_engine(golden, "ln")
    try:
        w = ovit.perturb_layernorms(ovit.make_visual_weights(1234))
        tiles = u8_to_tiles(torch.from_numpy(g["u8"]).cuda(), e.op_torch_dtype)
        x = e.encode_score(tiles, stop_after_layer=11)["x"].cpu()
        res = e.encode_score(tiles, want_logits=True)
        y = torch.nn.functional.layer_norm(x[:, 0], (768,), w["ln_post.weight"], w["ln_post.bias"], 1e-5)
        probs, logits, fn = ovit.score(y @ w["proj"], golden["tables"]["text_features"])
        assert (res["feats"].cpu() - fn).abs().max().item() <= 2e-5
        assert (res["logits"].cpu() - logits).abs().max().item() <= 2e-3
        assert (res["probs"].cpu() - probs).abs().max().item() <= 1e-4
        margin = torch.sort(logits, dim=1).values
        clear = (margin[:, -1] - margin[:, -2]) > 5e-3
        assert torch.equal(res["top1"].cpu()[clear].long(), logits.argmax(dim=1)[clear])
    finally:
        e.close()


@pytest.mark.parametrize("V", [4, 6, 10])
def test_gpu_vote_matches_reference(golden, V):
    from vilgod_b200.engine import CLASS_LIST, Engine
    g = golden["vote"]
    e = Engine(num_views=V)
    try:
        e.set_text_features(torch.randn(24, 512))
        idx, scores = g[f"idx{V}"], g[f"scores{V}"]
        C = idx.shape[0]
        probs = torch.zeros(C * V, 24)
        probs[torch.arange(C * V), torch.from_numpy(idx.reshape(-1)).long()] = torch.from_numpy(scores.reshape(-1))
        vc, vs = e.vote(probs.cuda(), torch.from_numpy(idx.reshape(-1)).int().cuda())
        names = np.asarray(e.mapped_names)[vc.cpu().numpy()]
        assert np.array_equal(names, g[f"voted_name{V}"])
        assert np.array_equal(vs.cpu().numpy(), g[f"voted_score{V}"])
    finally:
        e.close()


def test_error_paths(golden):
    from vilgod_b200 import _lib
    from vilgod_b200.engine import Engine, VilgodError
    e = Engine(num_views=4)
    try:
        tiles = torch.zeros(2, 196, 256, dtype=e.op_torch_dtype, device="cuda")
        e.num_prompts = 24
        with pytest.raises(VilgodError) as ei:
            e.encode_score(tiles)
        assert ei.value.code == _lib.VG_ESTATE
        with pytest.raises(ValueError):
            e.set_text_features(torch.randn(24, 100))
        with pytest.raises(KeyError):
            e.load_vit_weights({"conv1.weight": torch.zeros(768, 3, 16, 16)})
    finally:
        e.close()
    with pytest.raises(VilgodError):
        Engine(num_views=4, resolution=160)
