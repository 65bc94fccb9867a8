"""GPU: mmdti_b200.optim.FusedAdam reproduces torch.optim.Adam (the reference's optimizer, tasks/trainer.py:160-162)
step for step, keeps the bf16 weight shadows in sync, and survives CUDA-graph replay (device-resident step count)."""
import pytest
import torch

from conftest import rel_err
from mmdti_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1536, 512), (1536,), (7,), (31, 512), (1, 128), (961, 1), (3,), (2048, 512)]
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.05).cuda()) for s in shapes]


def test_fused_adam_matches_torch_adam():
    pa, pb = _params(1), _params(1)
    ref = torch.optim.Adam(pa, lr=1e-3, eps=1e-6)
    shadows = {pb[0]: pb[0].detach().bfloat16(), pb[7]: pb[7].detach().bfloat16()}
    mine = FusedAdam(pb, lr=1e-3, eps=1e-6, shadows=shadows)
    g = torch.Generator().manual_seed(2)
    for step in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.1 + step)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        mine.step()
        for a, b in zip(pa, pb):
            assert rel_err(b, a) < 2e-6, (step, a.shape)
    for p, sh in shadows.items():
        assert torch.equal(sh, p.detach().bfloat16())
    m, v = mine.state_for(pb[3])
    st = ref.state[pa[3]]
    assert rel_err(m, st["exp_avg"]) < 1e-6 and rel_err(v, st["exp_avg_sq"]) < 1e-6


def test_fused_adam_under_cuda_graph():
    pa, pb = _params(3), _params(3)
    ref = torch.optim.Adam(pa, lr=1e-3, eps=1e-6)
    mine = FusedAdam(pb, lr=1e-3, eps=1e-6)
    gen = torch.Generator(device="cuda").manual_seed(11)
    grads = [torch.randn(p.shape, device="cuda", generator=gen) for p in pb]
    for p, g in zip(pb, grads):
        p.grad = g                              # static gradient buffers, as in a captured training step
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        mine.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        mine.step()
    for _ in range(3):
        graph.replay()
    for _ in range(4):                          # warm-up step + 3 replays (the capture itself executes nothing)
        for a, g in zip(pa, grads):
            a.grad = g.clone()
        ref.step()
    torch.cuda.synchronize()
    assert int(mine.step_count.item()) == 4
    for a, b in zip(pa, pb):
        assert (b - a).abs().max().item() < 1e-6 + 1e-5 * a.abs().max().item()      # lr = 1e-3: updates are O(1e-3)


def test_fused_adam_reads_external_gradient_buffers():
    """grad_source + grad_scale: the flat all-reduced buckets of dist.OverlappedGradReducer(keep_flat=True) are consumed
    in place (summed gradients, scaled by 1/world inside the kernel); p.grad is ignored"""
    pa, pb = _params(5), _params(5)
    world = 4
    ref = torch.optim.Adam(pa, lr=1e-3, eps=1e-6)
    offs, off = [], 0
    for p in pb:
        offs.append(off)
        off += (p.numel() + 3) // 4 * 4
    flat = torch.zeros(off, device="cuda")
    slot = {p: flat[o:o + p.numel()].view_as(p) for p, o in zip(pb, offs)}
    mine = FusedAdam(pb, lr=1e-3, eps=1e-6, grad_scale=1.0 / world, grad_source=lambda p: slot[p])
    g = torch.Generator().manual_seed(6)
    for step in range(3):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad = gr.clone()
            slot[b].copy_(gr * world)                # what an all-reduce(SUM) over `world` identical ranks leaves
            b.grad = None
        ref.step()
        mine.step()
        for a, b in zip(pa, pb):
            assert rel_err(b, a) < 2e-6, (step, a.shape)


def test_graphed_step_prefetch_pipeline():
    """GraphedStep.prefetch: batches staged on a copy stream are consumed in order, results equal direct calls"""
    from mmdti_b200.graph import GraphedStep
    w = torch.randn(16, 8, device="cuda")

    def fn(a, b):
        return (a @ w).sum() + b.float().sum()

    ex = [torch.zeros(4, 16, device="cuda"), torch.zeros(4, dtype=torch.int64, device="cuda")]
    gs = GraphedStep(fn, ex, warmup=1)
    batches = [(torch.randn(4, 16).pin_memory(), torch.randint(0, 9, (4,)).pin_memory()) for _ in range(4)]
    want = [float(fn(a.cuda(), b.cuda())) for a, b in batches]
    gs.prefetch(*batches[0])
    got = []
    for i in range(4):
        out = gs()
        if i + 1 < 4:
            gs.prefetch(*batches[i + 1])
        got.append(float(out.item()))
    assert got == pytest.approx(want, rel=1e-5)
    assert float(gs(*batches[2]).item()) == pytest.approx(want[2], rel=1e-5)      # direct call still works
    with pytest.raises(Exception):
        gs()
    gs.close()


def test_graphed_training_step_matches_eager(report):
    """N replays of a graphed forward + backward + FusedAdam step == the same number of eager steps with zero_grad
    between them (dropout off).  Guards against the captured backward recording ``grad += new`` onto gradients the
    warm-up left defined: every replay would then hand the optimizer the SUM of all previous steps' gradients.
    Run with lr = 0 so that the parameters stay put and the comparison is exact up to the summation order of the atomics
    (with lr > 0 Adam turns the round-off noise of analytically-zero gradients -- key bias, per-head pair-bias constant
    -- into +-lr steps, and the two runs drift apart for reasons unrelated to the graph): the optimizer's first and
    second moments then are pure functions of the gradient history, (1 - beta^n) g and (1 - beta2^n) g^2."""
    import mmdti_b200
    from mmdti_b200.data import synthetic_molecules
    from mmdti_b200.graph import GraphedStep
    from mmdti_b200.models.encoder import UnimolEncoder
    dev = "cuda"
    tokens, dist, et, _ = synthetic_molecules(4, 14, seed=3)
    inputs = [tokens.to(dev), dist.to(dev), et.to(dev)]
    g = torch.randn(4, 16, 512, generator=torch.Generator().manual_seed(1)).to(dev)
    n_warm, n_steps = 3, 4
    for act, pair in (("fp32", "fp32"), ("bf16", "bf16")):
        out = []
        for graphed in (False, True):
            torch.manual_seed(0)
            m = UnimolEncoder(encoder_layers=2).to(dev).eval()
            opt = FusedAdam(m.parameters(), lr=0.0, eps=1e-6, shadows=m.encoder.use_external_lowp())

            def step(*inp):
                loss = (m(*inp).float() * g).sum()
                loss.backward()
                opt.step()
                return loss.detach()

            with mmdti_b200.precision(act=act, pair=pair):
                if graphed:
                    # the warm-up steps ARE executed optimizer steps; the capture pass itself executes nothing
                    gs = GraphedStep(step, inputs, warmup=n_warm, params=list(m.parameters()))
                    for _ in range(n_steps):
                        gs(*inputs)
                    gs.close()
                else:
                    for _ in range(n_warm + n_steps):
                        opt.zero_grad(set_to_none=True)
                        step(*inputs)
            torch.cuda.synchronize()
            assert int(opt.step_count.item()) == n_warm + n_steps
            out.append((torch.cat([p.grad.detach().reshape(-1) for p in m.parameters()]), opt.exp_avg.clone(), opt.exp_avg_sq.clone()))
        (ge, me, ve), (gg, mg, vg) = out
        e_g, e_m, e_v = rel_err(gg, ge), rel_err(mg, me), rel_err(vg, ve)
        report("graphed_vs_eager", act, "last grad=%.1e exp_avg=%.1e exp_avg_sq=%.1e" % (e_g, e_m, e_v))
        # an accumulating graph would be off by a factor of n_steps in the gradient and ~4x / ~16x in the moments
        assert e_g < 1e-4 and e_m < 1e-4 and e_v < 1e-4, (act, e_g, e_m, e_v)
