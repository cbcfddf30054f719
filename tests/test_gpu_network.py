"""Score-network parity: CUDA engine (fp16 tensor-core operands, fp32 accumulate) vs the oracle (plain PyTorch
fp32, TF32 off) on the same seeded inputs and non-degenerate weights; also against the committed reference
fixture.  Tolerance: 1e-3 relative L2 (BASELINE.json north_star)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def tol(net):
    # mixed (the default), fp16c8 and fp16x3 must meet the north-star 1e-3; the single-pass modes are opt-in speed
    # modes with TF32-class (11-bit) operands and are only required to stay in that class
    return TOL if net.precision in ("mixed", "fp16c8", "fp16x3") else 3e-3


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def sd():
    from oracle.weights import make_state_dict
    return make_state_dict(0)


@pytest.fixture(scope="module", params=["mixed", "fp16c8", "fp16x3", "fp16x2", "fp16"])
def net(sd, request):
    from buddy_b200.ncsnpp import NCSNppTime
    m = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2],
                   precision=request.param)
    m.load_state_dict(sd)
    return m.cuda().eval()


def test_state_dict_keys_match_reference(net):
    g = torch.load(os.path.join(GOLD, "state_dict_spec.pt"), weights_only=False)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == [(k, tuple(s)) for k, s in g["keys"]]


def test_forward_and_vjp_vs_reference_fixture(net):
    g = torch.load(os.path.join(GOLD, "net_small.pt"), weights_only=False)
    x = (randn(g["x_seed"], 2, 1, 8192) * g["x_scale"]).cuda().requires_grad_(True)
    cot = randn(g["cot_seed"], 2, 1, 8192).cuda()
    out = net(x, (0.25 * torch.log(g["sigma"])).cuda())
    (vjp,) = torch.autograd.grad((out * cot).sum(), x)
    e_out, e_vjp = rel(out.detach().cpu(), g["out"]), rel(vjp.cpu(), g["vjp"])
    print(f"\n[{net.precision}]", end="")
    print(f"\n[net small] fwd rel-L2 {e_out:.2e}  vjp rel-L2 {e_vjp:.2e}")
    assert e_out < tol(net) and e_vjp < tol(net)


def test_spectrogram_level_forward_vs_oracle(net, sd):
    from oracle import net as onet
    sdc = {k: v.cuda() for k, v in sd.items()}
    x = (randn(5, 1, 1, 8192) * 0.3).cuda()
    tc = torch.tensor([0.25 * torch.log(torch.tensor(0.05))]).cuda()
    spec = onet.net_stft(x)
    with torch.no_grad():
        ref = onet.ncsnpp_forward(sdc, spec, tc)
        got = super(type(net), net).forward(spec, tc)
    e = rel(torch.view_as_real(got), torch.view_as_real(ref))
    print(f"\n[{net.precision}]", end="")
    print(f"\n[net spec] fwd rel-L2 {e:.2e}")
    assert e < tol(net)


def test_full_size_forward_and_vjp_vs_oracle(net, sd):
    """One full 4 s utterance (65536 samples -> 256 x 528): forward and data-gradient vs the oracle on the GPU."""
    from oracle import net as onet
    sdc = {k: v.cuda() for k, v in sd.items()}
    x = (randn(6, 1, 1, 65536) * 0.2).cuda()
    tc = torch.tensor([0.25 * torch.log(torch.tensor(0.1))]).cuda()
    cot = randn(7, 1, 1, 65536).cuda() * 1e-3
    xr = x.clone().requires_grad_(True)
    ref = onet.ncsnpp_time_forward(sdc, xr, tc)
    (gref,) = torch.autograd.grad((ref * cot).sum(), xr)
    xg = x.clone().requires_grad_(True)
    out = net(xg, tc)
    (gout,) = torch.autograd.grad((out * cot).sum(), xg)
    e_out, e_vjp = rel(out.detach(), ref.detach()), rel(gout, gref)
    print(f"\n[{net.precision}]", end="")
    print(f"\n[net full] fwd rel-L2 {e_out:.2e}  vjp rel-L2 {e_vjp:.2e}")
    assert e_out < tol(net) and e_vjp < tol(net)


def test_other_lengths_forward_and_vjp_vs_oracle(sd):
    """The path is fully convolutional in time: a 6.1 s utterance (98304 samples -> 769 frames -> 784
    padded) and a ragged 1.3 s one (20517 samples) against the oracle on the GPU, default precision."""
    from buddy_b200.ncsnpp import NCSNppTime
    from oracle import net as onet
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    sdc = {k: v.cuda() for k, v in sd.items()}
    for n in (98304, 20517):
        x = (randn(8, 1, 1, n) * 0.2).cuda()
        tc = torch.tensor([0.25 * torch.log(torch.tensor(0.07))]).cuda()
        cot = randn(9, 1, 1, n).cuda() * 1e-3
        xr = x.clone().requires_grad_(True)
        ref = onet.ncsnpp_time_forward(sdc, xr, tc)
        (gref,) = torch.autograd.grad((ref * cot).sum(), xr)
        xg = x.clone().requires_grad_(True)
        out = net(xg, tc)
        (gout,) = torch.autograd.grad((out * cot).sum(), xg)
        e_out, e_vjp = rel(out.detach(), ref.detach()), rel(gout, gref)
        print(f"\n[net n={n}] fwd rel-L2 {e_out:.2e}  vjp rel-L2 {e_vjp:.2e}")
        assert e_out < TOL and e_vjp < TOL


def test_query_blocked_attention_matches_dense(sd):
    """Long utterances run the bottleneck attention over query blocks with recomputation in the backward pass (no
    N x N tensor is ever stored): forced here on a 0.5 s utterance (N = 320 tokens, blocks of 128 / 128 / 64) and
    compared with the dense path — the same softmax rows: the forward agrees exactly, the data-gradient to the fp16
    rounding of dQ/dK/dV (fp32 accumulation over blocks vs one GEMM) seen through the single-pass dgrad convolutions
    downstream (a different rounding decision in an fp16 operand is a 2^-11 relative change of that element)."""
    from buddy_b200.engine import Engine
    from buddy_b200.spectral import NetSTFT
    st = NetSTFT("cuda")
    x = (randn(20, 2, 8192) * 0.2).cuda()
    tc = torch.full((2,), -0.6, device="cuda")
    cot = randn(21, 2, 256, 80, 2).cuda()
    outs = []
    for blocked in (False, True):
        eng = Engine(sd, "cuda")
        if blocked:
            eng.ATTN_DENSE_BYTES, eng.ATTN_QBLOCK = 0, 128
        out, ctx = eng.forward(st.forward(x), tc, save=True)
        assert (ctx["attn"][3] is None) == blocked
        outs.append((out.clone(), eng.vjp(ctx, cot).clone()))
    e_f, e_b = rel(outs[1][0], outs[0][0]), rel(outs[1][1], outs[0][1])
    print(f"\n[blocked attention vs dense] fwd {e_f:.2e} vjp {e_b:.2e}")
    assert e_f < 1e-4 and e_b < 5e-4


def test_long_form_30s_forward_and_vjp_vs_oracle(sd):
    """BASELINE configs[4]: one 30 s utterance (480 000 samples -> 256 x 3760 spectrogram, attention over 15 040 tokens
    in query blocks, GroupNorm statistics over 7x the pixels) against the oracle on the GPU, default precision."""
    from buddy_b200.ncsnpp import NCSNppTime
    from oracle import net as onet
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    sdc = {k: v.cuda() for k, v in sd.items()}
    n = 480000
    x = (randn(30, 1, 1, n) * 0.2).cuda()
    tc = torch.tensor([0.25 * torch.log(torch.tensor(0.07))]).cuda()
    cot = randn(31, 1, 1, n).cuda() * 1e-3
    xr = x.clone().requires_grad_(True)
    ref = onet.ncsnpp_time_forward(sdc, xr, tc)
    (gref,) = torch.autograd.grad((ref * cot).sum(), xr)
    ref = ref.detach()
    torch.cuda.empty_cache()
    xg = x.clone().requires_grad_(True)
    out = net(xg, tc)
    assert net.engine()._attn_blocked(1, 15040)
    (gout,) = torch.autograd.grad((out * cot).sum(), xg)
    e_out, e_vjp = rel(out.detach(), ref), rel(gout, gref)
    print(f"\n[net 30 s] fwd rel-L2 {e_out:.2e}  vjp rel-L2 {e_vjp:.2e}")
    assert e_out < TOL and e_vjp < TOL


@pytest.mark.parametrize("n,B", [(2048, 3), (4000, 1), (131072, 2)])
def test_edge_lengths_forward_and_vjp_vs_oracle(sd, n, B):
    """Very short (0.13 s: 17 frames -> 32 padded, 4-frame bottleneck), ragged (4000 samples) and 8 s utterances, batched:
    forward and data-gradient vs the oracle on the GPU (tile shapes, stacked-tile and narrow-N heuristics all change
    with the size)."""
    from buddy_b200.ncsnpp import NCSNppTime
    from oracle import net as onet
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    sdc = {k: v.cuda() for k, v in sd.items()}
    x = (randn(40, B, 1, n) * 0.2).cuda()
    tc = (0.25 * torch.log(torch.linspace(0.02, 0.4, B))).cuda()
    cot = randn(41, B, 1, n).cuda() * 1e-3
    xr = x.clone().requires_grad_(True)
    ref = onet.ncsnpp_time_forward(sdc, xr, tc)
    (gref,) = torch.autograd.grad((ref * cot).sum(), xr)
    xg = x.clone().requires_grad_(True)
    out = net(xg, tc)
    (gout,) = torch.autograd.grad((out * cot).sum(), xg)
    for b in range(B):
        e_out, e_vjp = rel(out[b].detach(), ref[b].detach()), rel(gout[b], gref[b])
        print(f"\n[net n={n} utt {b}] fwd rel-L2 {e_out:.2e}  vjp rel-L2 {e_vjp:.2e}")
        assert e_out < TOL and e_vjp < TOL
