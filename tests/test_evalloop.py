"""Eval-loop parity.  CPU: the oracle port and the product loop reproduce the perplexity the LIVE
reference printed for the committed tiny-OPT fixture.  GPU: the packed model stays within 1e-3."""
import pytest
import torch

from conftest import load_tiny_opt


def test_oracle_port_reproduces_reference_ppl():
    from oracle.evalloop import reference_eval
    model, parts, ids, ref_ppl = load_tiny_opt()
    ppl = reference_eval(model, ids)
    assert abs(ppl - ref_ppl) / ref_ppl < 1e-6, (ppl, ref_ppl)


def test_product_loop_matches_reference_on_dense_model():
    from quip_b200.opt import opt_eval
    from quip_b200 import evalloop
    model, parts, ids, ref_ppl = load_tiny_opt()
    ppl = opt_eval(model, ids, torch.device('cpu'), verbose=False)
    assert abs(ppl - ref_ppl) / ref_ppl < 1e-6, (ppl, ref_ppl)
    # sample-major order gives the same NLLs as the reference's layer-major order
    S = model.seqlen
    nll = sum(float(evalloop.sample_nll(model, evalloop.OPT, ids[:, i * S:(i + 1) * S])) for i in range(ids.numel() // S))
    import math
    assert abs(math.exp(nll / ids.numel()) - ref_ppl) / ref_ppl < 1e-6
    # subset of samples (what a data-parallel rank evaluates)
    p01 = opt_eval(model, ids, torch.device('cpu'), sample_ids=[0, 1], verbose=False)
    p23 = opt_eval(model, ids, torch.device('cpu'), sample_ids=[2, 3], verbose=False)
    assert abs(math.exp((math.log(p01) + math.log(p23)) / 2) - ref_ppl) / ref_ppl < 1e-5


def test_pack_swaps_and_round_trips_through_load_quant():
    from quip_b200.opt import load_quant, opt_pack
    from quip_b200.quant import QuantLinear
    from quip_b200.modelutils import find_layers
    model, parts, ids, _ = load_tiny_opt()
    opt_pack(model, parts)
    q = find_layers(model, [QuantLinear])
    assert set(q) == set(parts) and len(q) == 12
    assert 'lm_head' not in q
    sd = model.state_dict()
    m2 = load_quant(model.config, sd)
    q2 = find_layers(m2, [QuantLinear])
    assert set(q2) == set(q)
    for k, v in sd.items():
        assert torch.equal(v, m2.state_dict()[k]), k
    # a dense fp16 state_dict written by the reference's --save still loads (as nn.Linear)
    dense, _, _, _ = load_tiny_opt()
    m3 = load_quant(dense.config, dense.state_dict())
    assert not find_layers(m3, [QuantLinear])


@pytest.mark.gpu
def test_packed_model_ppl_within_tolerance():
    from quip_b200.opt import opt_eval, opt_pack
    model, parts, ids, ref_ppl = load_tiny_opt()
    opt_pack(model, parts)
    dev = torch.device('cuda:0')
    ppl = opt_eval(model, ids, dev, verbose=False)
    assert abs(ppl - ref_ppl) / ref_ppl < 1e-3, (ppl, ref_ppl)      # north_star: ppl within 1e-3 of the reference
    import json, os
    from conftest import ROOT
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'parity_report.jsonl'), 'a') as f:
        f.write(json.dumps(dict(case='tiny_opt_2bit_incoh_ppl', ppl=ppl, reference_ppl=ref_ppl,
                                rel_delta=abs(ppl - ref_ppl) / ref_ppl)) + '\n')
