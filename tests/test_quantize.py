"""The producer side (quip_b200/quantize.py: calibration Hessian, random butterflies, LDLQ, LayerParts) on the CPU, against
the oracle's restatements of the reference (oracle/ldlq.py, oracle/quantflow.py, oracle/butterfly.py, oracle/forward.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import butterfly as obf
from oracle import forward as ofw
from oracle import ldlq as oldlq
from oracle import quantflow as oqf
from quip_b200 import quantize as qz


def _np_bpp(b):
    return ([b.B0.numpy(), b.B1.numpy()], b.p_in.numpy(), b.p_out.numpy())


@pytest.mark.parametrize('n,mode', [(96, 'blocked'), (128, 'kron'), (384, 'noperm'), (640, 'blocked')])
def test_butterfly_apply_matches_reference_structure(n, mode):
    g = torch.Generator().manual_seed(n)
    b = qz.random_butterfly(n, mode, g)
    x = torch.randn(n, 5, generator=g)
    want = obf.mul_butterfly(_np_bpp(b), x.numpy())                        # method.py:46-67 restated
    np.testing.assert_allclose(qz.butterfly_apply(b, x).numpy(), want, rtol=0, atol=3e-5)
    back = qz.butterfly_apply_t(b, qz.butterfly_apply(b, x))               # orthogonal: M^T M = I
    np.testing.assert_allclose(back.numpy(), x.numpy(), atol=3e-5)
    D = qz.butterfly_apply(b, torch.eye(n))
    np.testing.assert_allclose((D @ D.T).numpy(), np.eye(n), atol=3e-5)
    if mode == 'kron':
        assert b.B0.shape[0] == 1 and b.B1.shape[0] == 1
    if mode == 'noperm':
        assert torch.equal(b.p_in, torch.arange(n))


def test_blocked_ldlq_agrees_with_the_sequential_oracle():
    z = np.load(os.path.join(GOLDEN, 'ldlq.npz'))
    for case in z['cases']:
        key, nbits, npasses = str(case).split(':')
        ci, meth = key.split('_')
        w, H = torch.from_numpy(z[f'{ci}_w']), torch.from_numpy(z[f'{ci}_H'])
        fn = qz.ldlq_round if meth == 'ldlq' else qz.ldlq_rg_round
        for block in (32, 1024):
            got = fn(w, H, int(nbits), int(npasses), block=block)
            want = torch.from_numpy(z[f'{key}_out'])                      # the live reference's codes
            assert float((got == want).float().mean()) > 0.995, (case, block)
            loss = lambda q: float(torch.trace((q - w) @ H @ (q - w).T))
            assert loss(got) < 1.005 * loss(want), (case, block)


def test_hessian_accumulator_matches_reference_semantics():
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(2, 40, 64, generator=g).half(), torch.randn(1, 40, 64, generator=g).half()]
    acc = qz.HessianAccumulator(64)
    for x in xs:
        acc.add_batch(x)
    H, n = oqf.accumulate_hessian(xs)
    assert acc.batches == n == 3                                            # batches, not tokens (method.py:105,118)
    np.testing.assert_allclose(acc.result().numpy(), oqf.finalize_hessian(H, n).numpy(), rtol=1e-6)


@pytest.mark.parametrize('bits,method,qfn,incoh,rescale', [(2, 'ldlq', 'b', 'blocked', True), (2, 'ldlqRG', 'b', 'kron', True),
                                                           (4, 'ldlq', 'a', None, False), (3, 'ldlq', 'a', 'noperm', True)])
def test_quantize_linear_parts_are_consistent_and_better_than_nearest(bits, method, qfn, incoh, rescale):
    g = torch.Generator().manual_seed(bits * 7 + len(method))
    N, K = 96, 128
    W0 = (torch.randn(N, K, generator=g) * 0.05).half()
    W0[:, 3] *= 8                                                            # an outlier column: what incoherence spreads out
    X = (torch.randn(1, 512, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).half()
    acc = qz.HessianAccumulator(K)
    acc.add_batch(X)
    H = acc.result()
    bias = torch.randn(N, generator=g).half()
    parts, W_hat = qz.quantize_linear(W0, H, bits=bits, method=method, greedy_passes=1, qfn=qfn, rescale=rescale, incoh=incoh,
                                      bias=bias, generator=g, return_dense=True)
    assert parts.codes.dtype == torch.uint8 and int(parts.codes.max()) <= 2 ** bits - 1
    # the parts describe the same matrix as the dense fake-quantised weight (reference postproc, method.py:195-214)
    pn = dict(bits=bits, qfn=qfn, codes=parts.codes.numpy(), scales=parts.scales.numpy(), zeros=parts.zeros.numpy(),
              bias=bias.numpy(), scaleWH=None if parts.scaleWH is None else parts.scaleWH.numpy(),
              U=None if parts.U is None else _np_bpp(parts.U), V=None if parts.V is None else _np_bpp(parts.V))
    W_ref = ofw.w_ref(pn)
    assert np.mean(W_ref.view(np.uint16) == W_hat.numpy().view(np.uint16)) > 0.98
    np.testing.assert_allclose(W_ref.astype(np.float32), W_hat.float().numpy(), atol=2e-3 * float(W0.abs().max()))
    # ... and the kernel-side plan of those parts reproduces the dense forward within the path's tolerance
    x = (torch.randn(6, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).half().numpy()
    y_dense = ofw.dense_forward(x, W_ref, bias.numpy())
    y_kernel = ofw.kernel_forward(x, pn)
    assert ofw.rel_err(y_kernel, y_dense) < 1e-3
    # adaptive rounding beats nearest rounding on the same grid in the proxy loss
    loss = qz.proxy_loss(W_hat.float(), W0.float(), H)
    assert loss < {2: 0.35, 3: 0.08, 4: 0.02}[bits], loss


def test_quantized_parts_pack_on_the_host_and_round_trip():
    from quip_b200 import quant as Q
    g = torch.Generator().manual_seed(4)
    W0 = (torch.randn(64, 128, generator=g) * 0.05).half()
    acc = qz.HessianAccumulator(128)
    acc.add_batch((torch.randn(1, 300, 128, generator=g)).half())
    parts = qz.quantize_linear(W0, acc.result(), bits=2, qfn='b', incoh='blocked', generator=g)
    ql = Q.QuantLinear(infeatures=128, outfeatures=64, **Q.spec_from_parts(parts))
    ql.pack_parts(parts)
    assert ql.qweight.dtype == torch.int32 and ql.qweight.numel() == Q.packed_words(64, 128, 2)
    plan_codes = Q.unpack_codes(ql.qweight, 64, 128, 2)
    assert sorted(np.bincount(plan_codes.numpy().ravel(), minlength=4)) == sorted(np.bincount(parts.codes.numpy().ravel(), minlength=4))


def test_quantize_model_walks_the_decoder_stack():
    """Tiny random OPT on the CPU: every decoder Linear gets parts, the fake-quantised model stays close to the original, and
    the 4-bit quantisation is closer than the 2-bit one."""
    from transformers import OPTConfig, OPTForCausalLM
    from quip_b200 import evalloop
    cfg = OPTConfig(hidden_size=128, ffn_dim=256, num_hidden_layers=2, num_attention_heads=4, vocab_size=256,
                    max_position_embeddings=64, word_embed_proj_dim=128)
    torch.manual_seed(0)
    ref = OPTForCausalLM(cfg).float().eval()
    ref.seqlen = 32
    ids = [torch.randint(0, 256, (1, 32), generator=torch.Generator().manual_seed(i)) for i in range(6)]
    with torch.no_grad():
        base = ref(ids[0]).logits
    errs = {}
    for bits in (2, 4):
        import copy
        m = copy.deepcopy(ref)
        parts = qz.quantize_model(m, evalloop.OPT, ids, bits=bits, method='ldlq', qfn='b', rescale=True, incoh='blocked',
                                  generator=torch.Generator().manual_seed(1), pack=False)
        assert len(parts) == 2 * 6 and all(k.startswith('model.decoder.layers.') for k in parts)
        assert parts['model.decoder.layers.1.fc1'].codes.shape == (256, 128)
        with torch.no_grad():
            errs[bits] = float((m(ids[0]).logits - base).norm() / base.norm())
    assert errs[4] < errs[2] and errs[4] < 0.2, errs


def test_reference_flags_map_to_quantiser_options():
    from types import SimpleNamespace as NS
    o = qz.options_from_args(NS(quant='ldlqRG', wbits=2, npasses=2, incoh_processing=True, percdamp=0.01, qfn='a'))
    assert o == dict(bits=2, method='ldlq_rg', greedy_passes=2, qfn='b', rescale=True, incoh='blocked', percdamp=0.01, damp=True)
    o = qz.options_from_args(NS(quant='ldlq', wbits=4, npasses=0, pre_proj=True, pre_proj_extra=1, pre_gptqH=True, qfn='a'))
    assert o['incoh'] == 'kron' and not o['rescale'] and o['qfn'] == 'a' and o['damp']
    o = qz.options_from_args(NS(quant='nearest', wbits=3))
    assert o['method'] == 'nearest' and o['incoh'] is None and not o['damp']
    for bad in (NS(quant='gptq', wbits=4), NS(quant='ldlq', wbits=4, unbiased=True), NS(quant='allbal', wbits=2)):
        with pytest.raises(NotImplementedError):
            qz.options_from_args(bad)


def test_nearest_is_plain_rounding_and_ldlq_beats_it():
    g = torch.Generator().manual_seed(9)
    W0 = (torch.randn(32, 128, generator=g) * 0.05).half()
    acc = qz.HessianAccumulator(128)
    acc.add_batch((torch.randn(1, 400, 128, generator=g) * (1 + 3 * torch.rand(128, generator=g))).half())
    H = acc.result()
    near, wn = qz.quantize_linear(W0, H, bits=3, method='nearest', qfn='a', rescale=False, incoh=None, damp=False, return_dense=True)
    x = W0.float()
    lo, hi = torch.minimum(x.min(1)[0], torch.zeros(32)), torch.maximum(x.max(1)[0], torch.zeros(32))
    scale = ((hi - lo) / 7).reshape(-1, 1)
    zero = torch.round(-lo.reshape(-1, 1) / scale)
    assert torch.equal(near.codes, torch.clamp(torch.round(x / scale) + zero, 0, 7).to(torch.uint8))   # quant.py:6-9
    ldl, wl = qz.quantize_linear(W0, H, bits=3, method='ldlq', qfn='a', rescale=False, incoh=None, return_dense=True)
    assert qz.proxy_loss(wl.float(), W0.float(), H) < qz.proxy_loss(wn.float(), W0.float(), H)


def test_opt_sequential_pack_save_and_reload_round_trip(tmp_path):
    """Flags -> opt_sequential -> opt_pack -> state_dict -> load_quant: the packed checkpoint carries everything."""
    from types import SimpleNamespace as NS
    from transformers import OPTConfig, OPTForCausalLM
    from quip_b200 import opt as O
    from quip_b200.quant import QuantLinear
    cfg = OPTConfig(hidden_size=128, ffn_dim=256, num_hidden_layers=1, num_attention_heads=4, vocab_size=256,
                    max_position_embeddings=64, word_embed_proj_dim=128)
    torch.manual_seed(0)
    m = OPTForCausalLM(cfg).float().eval()
    m.seqlen = 32
    loader = [(torch.randint(0, 256, (1, 32), generator=torch.Generator().manual_seed(i)), None) for i in range(4)]
    args = NS(quant='ldlq', wbits=2, npasses=0, incoh_processing=True, percdamp=0.01)
    parts = O.opt_sequential(m, loader, 'cpu', args, generator=torch.Generator().manual_seed(2))
    assert set(parts) == {f'model.decoder.layers.0.{n}' for n in
                          ('self_attn.q_proj', 'self_attn.k_proj', 'self_attn.v_proj', 'self_attn.out_proj', 'fc1', 'fc2')}
    O.opt_pack(m.half(), parts)
    sd = m.state_dict()
    path = tmp_path / 'packed.pt'
    torch.save(sd, path)
    m2 = O.load_quant(cfg, str(path))
    q1 = {n: mod for n, mod in m.named_modules() if isinstance(mod, QuantLinear)}
    q2 = {n: mod for n, mod in m2.named_modules() if isinstance(mod, QuantLinear)}
    assert set(q1) == set(q2) == set(parts)
    for n in q1:
        assert q2[n].bits == 2 and q2[n].incoh == 'blocked' and q2[n].rescale
        assert torch.equal(q1[n].qweight, q2[n].qweight)
    sd2 = m2.state_dict()
    assert set(sd) == set(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)


# ---- the product quantiser against the LIVE REFERENCE's outputs (tests/golden, written by oracle/gen_golden_*.py) ----
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_product_rounding_reproduces_reference_codes():
    """quip_b200.quantize.ldlq_round / ldlq_rg_round on the (w, H) of tests/golden/ldlq.npz: with one block (the
    reference's column order of additions) the codes are the reference's bit for bit; 32-column blocks regroup fp32 sums
    and may flip a near-tie."""
    z = np.load(os.path.join(GOLDEN_DIR, 'ldlq.npz'))
    for case in z['cases']:
        key, nbits, npasses = str(case).split(':')
        ci, meth = key.split('_')
        w, H = torch.from_numpy(z[f'{ci}_w']), torch.from_numpy(z[f'{ci}_H'])
        fn = qz.ldlq_round if meth == 'ldlq' else qz.ldlq_rg_round
        want = z[f'{key}_out']
        got = fn(w, H, int(nbits), int(npasses), block=w.shape[1]).numpy()
        np.testing.assert_array_equal(got, want, err_msg=str(case))
        blocked = fn(w, H, int(nbits), int(npasses), block=32).numpy()
        assert np.mean(blocked == want) > 0.995, case


@pytest.mark.parametrize('name', ['plain_a', 'rescale_a'])
def test_product_quantize_linear_reproduces_reference_layers(name):
    """Raw fp16 weight + calibration activations -> codes, per-row grid, 1/s: the reference's own quantised layer
    (Balance.preproc -> fasterquant, captured in tests/golden/quantflow_*.npz)."""
    z = np.load(os.path.join(GOLDEN_DIR, f'quantflow_{name}.npz'))
    W0, X = torch.from_numpy(z['W0']), torch.from_numpy(z['X'])
    acc = qz.HessianAccumulator(W0.shape[1])
    acc.add_batch(X.unsqueeze(0))
    method = {'ldlqRG': 'ldlq_rg'}.get(str(z['method']), str(z['method']))
    parts = qz.quantize_linear(W0, acc.result(), bits=int(z['bits']), method=method, greedy_passes=int(z['npasses']),
                               qfn=str(z['qfn']), rescale=bool(int(z['rescale'])), incoh=None)
    assert np.mean(parts.codes.numpy() == z['codes']) > 0.995
    assert np.mean(parts.grid.numpy().view(np.uint16) == z['grid'].view(np.uint16)) > 0.995
    if int(z['rescale']):
        np.testing.assert_allclose(parts.scaleWH.numpy(), z['scaleWH'], rtol=1e-5)
    # per-row grid: W = scales * code - zeros with scales = scale, zeros = zero * scale (quant.py:186-191)
    np.testing.assert_allclose(parts.scales.numpy().ravel(), z['scale'].ravel(), rtol=1e-6)
    np.testing.assert_allclose(parts.zeros.numpy().ravel(), (z['zero'] * z['scale']).ravel(), rtol=1e-6)


def test_optq_equals_ldlq_on_a_fake_layer():
    """The reference's one self-check (optq_ldlq_equiv.py, which needs CUDA and only prints the agreement): OPTQ, restated
    in oracle/optq.py, and the product's LDLQ pick the same grid points -- LDLQ walks columns last to first, OPTQ first to
    last, so LDLQ runs on the column-reversed problem.  Both must beat nearest rounding on the proxy loss."""
    from oracle import optq
    bits = 3
    w, H = optq.fake_layer(192, 160, seed=1)
    codes_optq, scale, zero = optq.optq_codes(w, H, bits)
    t = torch.clamp(w / scale + zero, 0, 2 ** bits - 1)                  # the weight in grid units
    rev = torch.arange(w.shape[1] - 1, -1, -1)
    codes_ldlq = qz.ldlq_round(t[:, rev].float(), H[rev][:, rev].float(), bits, 0, block=64).double()[:, rev]
    # fp32 feedback sums and the tie rule (floor(v + 0.5) vs round-half-even) flip a near-tie now and then; the flip then
    # propagates along its row
    assert float((codes_ldlq == codes_optq).double().mean()) > 0.995
    deq = lambda c: scale * (c - zero)                                    # noqa: E731
    nearest = torch.clamp(torch.round(t), 0, 2 ** bits - 1)
    loss = lambda q: float(torch.trace((q - w) @ H @ (q - w).T))         # noqa: E731
    assert loss(deq(codes_ldlq)) < 0.7 * loss(deq(nearest)) and loss(deq(codes_optq)) < 0.7 * loss(deq(nearest))
    assert abs(loss(deq(codes_ldlq)) / loss(deq(codes_optq)) - 1) < 0.01
