"""BASELINE.json config cases other than the bench workload, as parity tests.

config[0]: OPT-125m shapes, 4-bit, qfn 'a' (no incoherence), 128-token sequence -- the reference contract
`pack(linear, scales, zeros)` (quant.py:185-191) + `opt_pack3`-style swap, plumbing on CPU and parity on GPU.
"""
import math

import pytest
import torch
import torch.nn as nn


def _opt125m_like(layers):
    from transformers import OPTConfig
    from quip_b200.opt import get_opt
    from quip_b200.synth import OPT_125M
    torch.manual_seed(0)
    cfg = OPTConfig(**{**OPT_125M, 'num_hidden_layers': layers, 'vocab_size': 2048})
    model = get_opt(cfg)
    model.seqlen = 128
    return model


def _fake_quantize_(model, bits):
    """What the reference's `--quant nearest --wbits 4` leaves behind: grid weights + per-layer quantizers."""
    from quip_b200.modelutils import find_layers
    from quip_b200.quant import Quantizer
    quantizers = {}
    for i, layer in enumerate(model.model.decoder.layers):
        for name, lin in find_layers(layer).items():
            q = Quantizer()
            q.configure(bits, perchannel=True, sym=False, qfn='a', mse=False)
            q.find_params(lin.weight.data, weight=True)
            lin.weight.data = q.quantize(lin.weight.data).to(lin.weight.dtype)
            quantizers[f'model.decoder.layers.{i}.{name}'] = q
    return quantizers


def _pack(model, quantizers, bits):
    from quip_b200.modelutils import find_layers
    from quip_b200.quant import QuantLinear, make_quant
    layers = {n: m for n, m in find_layers(model).items() if n in quantizers}
    make_quant(model, list(quantizers), bits=bits)
    q = find_layers(model, [QuantLinear])
    for name in q:
        q[name].pack(layers[name], quantizers[name].scale, quantizers[name].zero)
    return q


def test_config0_plumbing_on_cpu():
    from quip_b200.opt import load_quant
    model = _opt125m_like(2)
    quantizers = _fake_quantize_(model, 4)
    dense_sd = {k: v.clone() for k, v in model.state_dict().items()}
    q = _pack(model, quantizers, 4)
    assert len(q) == 12 and all(m.bits == 4 and m.incoh is None for m in q.values())
    # codes recovered from the grid weights are what the quantizer would emit
    name = 'model.decoder.layers.0.fc1'
    W = dense_sd[name + '.weight'].float()
    qz = quantizers[name]
    want = torch.round(W / qz.scale + qz.zero).to(torch.uint8)
    assert torch.equal(q[name].codes(), want)
    m2 = load_quant(model.config, model.state_dict())
    assert sum(isinstance(m, type(q[name])) for m in m2.modules()) == 12


@pytest.mark.gpu
def test_config0_opt125m_4bit_packed_matches_dense_fake_quant():
    from quip_b200.opt import opt_eval
    dev = torch.device('cuda:0')
    dense = _opt125m_like(12)
    quantizers = _fake_quantize_(dense, 4)
    ids = torch.randint(0, 2048, (1, 4 * 128), generator=torch.Generator().manual_seed(3))
    dense.to(dev)
    ppl_dense = opt_eval(dense, ids, dev, verbose=False)        # the reference's effective path: F.linear on grid weights
    packed = _opt125m_like(12)
    packed.load_state_dict({k: v.cpu() for k, v in dense.state_dict().items()})
    _pack(packed, quantizers, 4)
    packed.to(dev)
    ppl_packed = opt_eval(packed, ids, dev, verbose=False)
    assert math.isfinite(ppl_packed)
    assert abs(ppl_packed - ppl_dense) / ppl_dense < 2e-3, (ppl_packed, ppl_dense)
    # one layer, token by token vs all at once (the reference's M == 1 contract, quant.py:223)
    layer = packed.model.decoder.layers[0].fc1
    x = torch.randn(5, 768, device=dev).half()
    y_all = layer(x)
    y_one = torch.cat([layer(x[i:i + 1]) for i in range(5)])
    assert (y_all.float() - y_one.float()).norm() / y_all.float().norm() < 5e-4


@pytest.mark.gpu
def test_decode_benchmark_with_kv_cache():
    """benchmark() (opt.py:431-482): token-by-token decode through the packed M == 1 path, with --check ppl."""
    import sys, os
    from conftest import load_tiny_opt
    from quip_b200.opt import benchmark, opt_eval, opt_pack
    model, parts, ids, ref_ppl = load_tiny_opt()
    opt_pack(model, parts)
    dev = torch.device('cuda:0')
    model.to(dev)
    seq = ids[:, :model.seqlen]
    med, ppl_decode = benchmark(model, seq, check=True)
    ppl_eval = opt_eval(model, seq, dev, verbose=False)
    assert med > 0 and math.isfinite(ppl_decode)
    assert abs(ppl_decode - ppl_eval) / ppl_eval < 2e-2, (ppl_decode, ppl_eval)
