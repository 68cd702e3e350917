"""Flat fused optimizers (CPU path) against torch.optim: same update rules as the reference's Keras optimizers expose
(`getattr(tf.keras.optimizers, name)` in P2/01:154 - SGD / Adam / Adadelta), on ONE flat fp32 buffer."""
import pytest
import torch

from b200ddl import optim


def _run(ours, theirs_factory, steps=5, n=4096, seed=0):
    g = torch.Generator().manual_seed(seed)
    p0 = torch.randn(n, generator=g)
    p = p0.clone()
    grads = torch.zeros(n)
    w16 = torch.zeros(n, dtype=torch.bfloat16)
    ours.attach(p, grads, w16)
    ref = torch.nn.Parameter(p0.clone())
    theirs = theirs_factory([ref])
    for _ in range(steps):
        gr = torch.randn(n, generator=g)
        grads.copy_(gr)
        ours.begin_step()
        ours.step()
        ref.grad = gr.clone()
        theirs.step()
    return p, ref.detach(), w16


@pytest.mark.parametrize("momentum,nesterov,wd", [(0.0, False, 0.0), (0.9, False, 1e-4), (0.9, True, 1e-4)])
def test_sgd_matches_torch(momentum, nesterov, wd):
    p, ref, w16 = _run(optim.SGD(0.05, momentum=momentum, nesterov=nesterov, weight_decay=wd),
                       lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=momentum, nesterov=nesterov, weight_decay=wd))
    assert torch.allclose(p, ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(w16, p.to(torch.bfloat16))          # the bf16 working copy is the rounded fp32 master


def test_adam_matches_torch():
    p, ref, _ = _run(optim.Adam(1e-2, beta_1=0.9, beta_2=0.999, epsilon=1e-7, weight_decay=1e-3),
                     lambda ps: torch.optim.Adam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-7, weight_decay=1e-3))
    assert torch.allclose(p, ref, rtol=1e-4, atol=1e-6)


def test_adadelta_matches_torch():
    p, ref, _ = _run(optim.Adadelta(1.0, rho=0.95, epsilon=1e-6),
                     lambda ps: torch.optim.Adadelta(ps, lr=1.0, rho=0.95, eps=1e-6))
    assert torch.allclose(p, ref, rtol=1e-4, atol=1e-6)


def test_learning_rate_change_takes_effect_without_reattach():
    """Callbacks (warm-up, ReduceLROnPlateau) set `.lr`; the next begin_step() must pick it up (device hyper array)."""
    o = optim.SGD(0.1)
    p = torch.ones(8)
    g = torch.ones(8)
    o.attach(p, g)
    o.begin_step(); o.step()
    assert torch.allclose(p, torch.full((8,), 0.9))
    o.lr = 0.5
    o.begin_step(); o.step()
    assert torch.allclose(p, torch.full((8,), 0.4))


def test_state_dict_round_trip():
    a = optim.Adam(1e-2)
    pa, ga = torch.randn(16), torch.randn(16)
    a.attach(pa, ga)
    for _ in range(3):
        a.begin_step(); a.step()
    sd = a.state_dict()
    b = optim.Adam(1.0)
    pb, gb = pa.clone(), ga.clone()
    b.attach(pb, gb)
    b.load_state_dict(sd)
    a.begin_step(); a.step()
    b.begin_step(); b.step()
    assert b.t == a.t and torch.allclose(pa, pb)


def test_unknown_optimizer_name():
    with pytest.raises(ValueError):
        optim.get("Nadam")


def test_fused_update_owned_ranges_tile_every_bucket():
    """`DistributedOptimizer.owned_range` (used by `consolidate_state`) must reproduce the slicing of the fused
    all-reduce + SGD kernel: the ranks' ranges are disjoint, ordered, float4-aligned and cover the bucket's full vectors."""
    from b200ddl.parallel.dist_optimizer import DistributedOptimizer as D

    for lo, hi in [(0, 4096), (128, 128 + 250000), (64, 64 + 4 * 7), (0, 4 * 3)]:
        for world in (2, 3, 4, 8):
            rs = [D.owned_range(lo, hi, r, world) for r in range(world)]
            assert rs[0][0] == lo and rs[-1][1] == lo + (hi - lo) // 4 * 4
            for (a0, b0), (a1, b1) in zip(rs, rs[1:]):
                assert b0 == a1 and a0 <= b0 and (a0 - lo) % 4 == 0
            nvec = (hi - lo) // 4
            per = -(-nvec // world)
            assert all(b - a <= 4 * per for a, b in rs)
