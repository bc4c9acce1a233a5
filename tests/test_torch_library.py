"""The ``torch.library`` packaging of the fused TP+scatter (nequip_b200/torch_ops.py): what
``nequip-compile --modifiers enable_B200TensorProductScatter`` / train-time compile need at the reference's seam
(nequip/utils/fx.py:52-119, nequip/nn/compile.py:168-191, nequip/nn/grad_output.py:217-221).
CPU part: fake-tensor propagation, symbolic make_fx tracing (forward + first and second derivative graphs).
GPU part: opcheck, torch.compile == eager, second-order autograd against finite differences / fp64 gradgradcheck."""
import pytest
import torch
from torch.fx.experimental.proxy_tensor import make_fx

from nequip_b200 import known_signatures as ks
from nequip_b200 import ops, torch_ops
from nequip_b200.codegen import GenOptions


def _plan(i=0, layout="mul_ir"):
    sig = ks.reference_test_grid()[i]
    plan = ops.get_plan(sig.irreps_in1, sig.irreps_in2, sig.irreps_out, sig.instructions, GenOptions(layout=layout))
    return plan, torch_ops.register_plan(plan)


def test_fake_tensor_propagation_and_symbolic_trace():
    from torch._subclasses.fake_tensor import FakeTensorMode

    plan, key = _plan()
    with FakeTensorMode():
        x = torch.empty(8, plan.d_in, requires_grad=True)
        y = torch.empty(15, plan.s_dim, requires_grad=True)
        w = torch.empty(15, plan.weight_numel, requires_grad=True)
        dst, src = torch.empty(15, dtype=torch.long), torch.empty(15, dtype=torch.long)
        out = torch.ops.nequip_b200.tp_scatter(x, y, w, dst, src, key)
        assert out.shape == (8, plan.d_out)
        gx, gy, gw = torch.autograd.grad(out.sum(), [x, y, w], create_graph=True)
        assert gx.shape == x.shape and gy.shape == y.shape and gw.shape == w.shape and gx.requires_grad
        (ggw,) = torch.autograd.grad(gx.sum() + gy.sum(), [w])  # second order: a force loss differentiated w.r.t. weights
        assert ggw.shape == w.shape

    def energy_and_force_like(x, y, w, dst, src):
        out = torch.ops.nequip_b200.tp_scatter(x, y, w, dst, src, key)
        (gy,) = torch.autograd.grad(out.square().sum(), [y], create_graph=True)  # "forces"
        (gw,) = torch.autograd.grad(gy.square().sum(), [w])  # d(force loss)/d(weights)
        return out, gy, gw

    x = torch.randn(8, plan.d_in)
    y = torch.randn(15, plan.s_dim, requires_grad=True)
    w = torch.randn(15, plan.weight_numel, requires_grad=True)
    dst, src = torch.randint(0, 8, (15,)), torch.randint(0, 8, (15,))
    gm = make_fx(energy_and_force_like, tracing_mode="symbolic", _allow_non_fake_inputs=True)(x, y, w, dst, src)
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_function"]
    assert any("nequip_b200.tp_scatter.default" in t for t in targets)
    assert sum("nequip_b200.tp_scatter_bwd" in t for t in targets) >= 2  # first derivative + its derivative
    # no shape specialisation: N and E stay symbolic
    code = gm.code
    assert "15" not in code.replace("15,", "") or "sym" in code


@pytest.mark.gpu
def test_opcheck_and_compile_match_eager():
    plan, key = _plan(4)
    g = torch.Generator().manual_seed(0)
    N, E = 8, 15
    x = torch.randn(N, plan.d_in, generator=g).cuda().requires_grad_(True)
    y = torch.randn(E, plan.s_dim, generator=g).cuda().requires_grad_(True)
    w = torch.randn(E, plan.weight_numel, generator=g).cuda().requires_grad_(True)
    dst = torch.sort(torch.randint(0, N, (E,), generator=g)).values.cuda()
    src = torch.randint(0, N, (E,), generator=g).cuda()
    torch.library.opcheck(torch.ops.nequip_b200.tp_scatter.default, (x, y, w, dst, src, key),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))

    def f(x, y, w, dst, src):
        out = torch.ops.nequip_b200.tp_scatter(x, y, w, dst, src, key)
        return out.sin().sum()

    ref = f(x, y, w, dst, src)
    gref = torch.autograd.grad(ref, [x, y, w])
    cf = torch.compile(f, fullgraph=True, dynamic=True)
    got = cf(x, y, w, dst, src)
    ggot = torch.autograd.grad(got, [x, y, w])
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6)
    for a, b in zip(ggot, gref):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["mul_ir", "ir_mul"])
def test_second_order_autograd(layout):
    """create_graph=True through the op (training with a force loss): fp64 gradcheck + gradgradcheck."""
    plan, key = _plan(1, layout)
    g = torch.Generator().manual_seed(1)
    N, E = 5, 9
    x = torch.randn(N, plan.d_in, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    y = torch.randn(E, plan.s_dim, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    w = torch.randn(E, plan.weight_numel, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    dst = torch.sort(torch.randint(0, N, (E,), generator=g)).values.cuda()
    src = torch.randint(0, N, (E,), generator=g).cuda()
    fn = lambda x, y, w: torch.ops.nequip_b200.tp_scatter(x, y, w, dst, src, key)
    assert torch.autograd.gradcheck(fn, (x, y, w), eps=1e-6, atol=1e-7, rtol=1e-6, nondet_tol=1e-12)
    assert torch.autograd.gradgradcheck(fn, (x, y, w), eps=1e-6, atol=1e-7, rtol=1e-6, nondet_tol=1e-12)


@pytest.mark.gpu
def test_module_forward_goes_through_the_registered_op():
    from nequip_b200.nn.tp_scatter import B200TensorProductScatter

    sig = ks.reference_test_grid()[2]
    m = B200TensorProductScatter(sig.irreps_in1, sig.irreps_in2, sig.irreps_out, sig.instructions)
    N, E = 6, 11
    x = torch.randn(N, sig.d_in).cuda()
    y, w = torch.randn(E, sig.s_dim).cuda(), torch.randn(E, sig.weight_numel).cuda().requires_grad_(True)
    dst, src = torch.sort(torch.randint(0, N, (E,))).values.cuda(), torch.randint(0, N, (E,)).cuda()
    gm = make_fx(lambda x, y, w, d, s: m(x, y, w, d, s), tracing_mode="symbolic", _allow_non_fake_inputs=True)(x, y, w, dst, src)
    assert any("nequip_b200.tp_scatter" in str(n.target) for n in gm.graph.nodes)
    torch.testing.assert_close(gm(x, y, w, dst, src), m(x, y, w, dst, src))
