"""Graph-captured Llama decode (quip_b200/decode.py) against the eager HF forward with a KV cache -- the loop the
reference's benchmark() times (opt.py:431-482)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_llama(kv_heads):
    from transformers import LlamaConfig
    from quip_b200.synth import build_synthetic_model
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=kv_heads, vocab_size=320, max_position_embeddings=128)
    return build_synthetic_model(cfg, torch.device('cuda:0'), bits=2, incoh='blocked', rescale=True, seed=5, seqlen=64)


@pytest.mark.parametrize('kv_heads', [4, 2])
def test_graph_decoder_matches_eager_hf_decode(kv_heads):
    from quip_b200.decode import GraphDecoder
    model = _tiny_llama(kv_heads)
    ids = torch.randint(0, 320, (1, 14), generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        past, want = None, []
        for i in range(ids.shape[1]):
            out = model(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            past = out.past_key_values
            want.append(out.logits[0, -1].float())
        dec = GraphDecoder(model, max_len=32)
        eager = [dec.step(ids[0, i:i + 1])[0].float().clone() for i in range(ids.shape[1])]
        dec.reset()
        dec.capture()
        graph = [dec.step(ids[0, i:i + 1])[0].float().clone() for i in range(ids.shape[1])]
        # control: the HF decode against ITSELF with one-ulp flips in 3e-5 of its norm outputs.  The decoder restates the
        # fp16 glue (another attention kernel over the static cache, another order of the same roundings), and a random-init
        # model amplifies any such noise: the graph decode has to stay within 3x of what the HF loop does to itself.
        import bench
        norms = [m for layer in model.model.layers for m in (layer.input_layernorm, layer.post_attention_layernorm)] + [model.model.norm]
        past, ctrl = None, []
        for i in range(ids.shape[1]):
            hooks = bench._ulp_flip_hooks(norms, 3e-5, seed=i)
            try:
                out = model(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            finally:
                for hk in hooks:
                    hk.remove()
            past = out.past_key_values
            ctrl.append(out.logits[0, -1].float())
    worst = control = 0.0
    for i, (w, e, g, c) in enumerate(zip(want, eager, graph, ctrl)):
        assert torch.equal(e, g), i                                   # replay == the recorded computation
        worst = max(worst, float((g - w).norm() / w.norm()))
        control = max(control, float((c - w).norm() / w.norm()))
    assert worst < max(2e-3, 3.0 * control), (worst, control)         # round 1 accepted 2e-2 outright
    with pytest.raises(ValueError):
        for _ in range(40):
            dec.step(ids[0, :1])


def test_graph_decode_benchmark_reports_time_and_ppl():
    from quip_b200.decode import graph_decode_benchmark
    from quip_b200.evalloop import decode_benchmark
    model = _tiny_llama(4)
    ids = torch.randint(0, 320, (1, 12), generator=torch.Generator().manual_seed(4))
    sec, ppl = graph_decode_benchmark(model, ids, check=True)
    sec0, ppl0 = decode_benchmark(model, ids, check=True)
    assert sec > 0 and abs(ppl - ppl0) / ppl0 < 2e-2


def test_graph_decoder_batch_rows_are_independent_sequences():
    from quip_b200.decode import GraphDecoder
    model = _tiny_llama(4)
    ids = torch.randint(0, 320, (3, 10), generator=torch.Generator().manual_seed(8)).cuda()
    with torch.no_grad():
        batched = GraphDecoder(model, max_len=16, batch=3).capture()
        got = [batched.step(ids[:, i]).float().clone() for i in range(ids.shape[1])]
        for b in range(3):
            single = GraphDecoder(model, max_len=16, batch=1).capture()
            for i in range(ids.shape[1]):
                want = single.step(ids[b, i:i + 1])[0].float()
                err = float((got[i][b] - want).norm() / want.norm())
                assert err < 5e-3, (b, i, err)        # 3 tokens take the int8 two-tile route, 1 token the one-tile route


def test_graphed_sample_nll_equals_eager_and_feeds_eval_ppl():
    """The decoder stack + head + loss of a fixed-length sample replayed from one CUDA graph (evalloop.GraphedSampleNLL),
    with the q/k/v and gate/up sibling groups on their side streams inside the capture."""
    from quip_b200 import evalloop
    from quip_b200.quant import group_siblings
    model = _tiny_llama(4)
    group_siblings(model)
    ids = torch.randint(0, 320, (4, 1, 64), generator=torch.Generator().manual_seed(9)).cuda()
    with torch.no_grad():
        eager = [float(evalloop.sample_nll(model, evalloop.LLAMA, ids[i])) for i in range(4)]
        stepper = evalloop.enable_graphed_eval(model, evalloop.LLAMA, ids[0])
        graphed = [float(stepper(ids[i])) for i in range(4)]
    for e, g in zip(eager, graphed):
        assert abs(e - g) <= 1e-4 * abs(e), (e, g)
    flat = ids.reshape(1, -1)
    ppl_graph = evalloop.eval_ppl(model, evalloop.LLAMA, flat, torch.device('cuda:0'), verbose=False)
    model._quip_graph_step = None
    ppl_eager = evalloop.eval_ppl(model, evalloop.LLAMA, flat, torch.device('cuda:0'), verbose=False)
    assert abs(ppl_graph - ppl_eager) <= 1e-3 * ppl_eager
    with pytest.raises(ValueError):
        stepper(ids[0][:, :32])


def test_opt_graph_decoder_matches_eager_hf_decode():
    """The OPT step of GraphDecoder (packed q/k/v/out_proj/fc1/fc2 with bias, LayerNorm, learned positions) against the HF
    forward with a KV cache -- the reference's benchmark() loop (opt.py:431-482); graph replay == the eager step."""
    from transformers import OPTConfig
    import bench
    from quip_b200.decode import GraphDecoder
    from quip_b200.opt import benchmark
    from quip_b200.synth import build_synthetic_model
    cfg = OPTConfig(hidden_size=256, ffn_dim=1024, num_hidden_layers=2, num_attention_heads=4, vocab_size=320,
                    max_position_embeddings=128, word_embed_proj_dim=256)
    model = build_synthetic_model(cfg, torch.device('cuda:0'), bits=2, incoh='blocked', rescale=True, seed=6, seqlen=64)
    ids = torch.randint(0, 320, (1, 14), generator=torch.Generator().manual_seed(3)).cuda()
    dec_mod = model.model.decoder
    norms = [m for layer in dec_mod.layers for m in (layer.self_attn_layer_norm, layer.final_layer_norm)] + [dec_mod.final_layer_norm]

    def hf_decode(flip):
        past, res = None, []
        for i in range(ids.shape[1]):
            hooks = bench._ulp_flip_hooks(norms, 3e-5, seed=i) if flip else []
            try:
                out = model(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            finally:
                for hk in hooks:
                    hk.remove()
            past = out.past_key_values
            res.append(out.logits[0, -1].float())
        return res
    with torch.no_grad():
        want, ctrl = hf_decode(False), hf_decode(True)
        dec = GraphDecoder(model, max_len=32)
        eager = [dec.step(ids[0, i:i + 1])[0].float().clone() for i in range(ids.shape[1])]
        dec.reset()
        dec.capture()
        graph = [dec.step(ids[0, i:i + 1])[0].float().clone() for i in range(ids.shape[1])]
    worst = control = 0.0
    for i, (w, e, g, c) in enumerate(zip(want, eager, graph, ctrl)):
        assert torch.equal(e, g), i
        worst = max(worst, float((g - w).norm() / w.norm()))
        control = max(control, float((c - w).norm() / w.norm()))
    assert worst < max(2e-3, 3.0 * control), (worst, control)
    sec, ppl = benchmark(model, ids.cpu(), check=True, graph=True)
    sec0, ppl0 = benchmark(model, ids.cpu(), check=True)
    assert sec > 0 and abs(ppl - ppl0) / ppl0 < 2e-2
