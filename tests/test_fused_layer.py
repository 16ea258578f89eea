"""CPU tests of the fused Llama stack's host logic (quip_b200/fused.py) with the glue ops injected from oracle/glue.py,
against the HF decoder layers themselves."""
import os

import pytest
import torch

from oracle.glue import TorchGlue
from quip_b200 import evalloop, fused


def _tiny(dtype, nkv=4, layers=3):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=layers, num_attention_heads=4,
                      num_key_value_heads=nkv, vocab_size=199, max_position_embeddings=64)
    torch.manual_seed(0)
    m = LlamaForCausalLM(cfg).to(dtype).eval()
    m.seqlen = 32
    return m


@pytest.mark.parametrize('dtype,nkv', [(torch.float32, 4), (torch.float32, 2), (torch.float16, 4)])
def test_stack_with_torch_glue_equals_hf_layers(dtype, nkv):
    m = _tiny(dtype, nkv)
    ids = torch.randint(0, 199, (1, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        h, kw = evalloop.layer_inputs(m, evalloop.LLAMA, ids)
        ref = h
        for layer in m.model.layers:
            ref = evalloop._call_layer(layer, ref, kw)
        got = fused.llama_stack(list(m.model.layers), h.clone(), kw, ops=TorchGlue())
    if dtype == torch.float32:
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    else:
        # same fp16 rounding points; the attention kernel sees other strides, nothing else differs
        assert float((got.float() - ref.float()).norm() / ref.float().norm()) < 2e-3
    assert got.shape == ref.shape


def test_stack_of_one_layer_and_empty_stack():
    m = _tiny(torch.float32, layers=1)
    ids = torch.randint(0, 199, (1, 16), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        h, kw = evalloop.layer_inputs(m, evalloop.LLAMA, ids)
        ref = evalloop._call_layer(m.model.layers[0], h, kw)
        got = fused.llama_stack([m.model.layers[0]], h.clone(), kw, ops=TorchGlue())
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
        assert torch.equal(fused.llama_stack([], h.clone(), kw, ops=TorchGlue()), h)


def test_torch_glue_matches_the_hf_modules_bit_for_bit():
    from transformers.models.llama import modeling_llama as M
    g = torch.Generator().manual_seed(3)
    ops = TorchGlue()
    for dtype in (torch.float16, torch.float32):
        x = torch.randn(1, 8, 64, generator=g).to(dtype)
        r = torch.randn(1, 8, 64, generator=g).to(dtype)
        norm = M.LlamaRMSNorm(64, eps=1e-5).to(dtype)
        norm.weight.data = torch.randn(64, generator=g).to(dtype)
        s, y = ops.rmsnorm(x, norm.weight, norm.variance_epsilon, residual=r)
        assert torch.equal(s, r + x) and torch.equal(y, norm(r + x))
        assert torch.equal(ops.rmsnorm(x, norm.weight, norm.variance_epsilon), norm(x))
        # rotary: token-major storage in place vs HF on (1, heads, S, hd)
        S, nq, nkv, hd = 8, 4, 2, 16
        q = torch.randn(1, S, nq * hd, generator=g).to(dtype)
        k = torch.randn(1, S, nkv * hd, generator=g).to(dtype)
        ang = torch.rand(S, hd // 2, generator=g) * 6
        cos, sin = torch.cat((ang.cos(), ang.cos()), -1).to(dtype), torch.cat((ang.sin(), ang.sin()), -1).to(dtype)
        qe, ke = M.apply_rotary_pos_emb(q.view(1, S, nq, hd).transpose(1, 2), k.view(1, S, nkv, hd).transpose(1, 2),
                                        cos[None], sin[None])
        q2, k2 = q.clone(), k.clone()
        ops.rope_(q2, k2, cos, sin, hd)
        assert torch.equal(q2.view(1, S, nq, hd).transpose(1, 2), qe) and torch.equal(k2.view(1, S, nkv, hd).transpose(1, 2), ke)
        gate, up = torch.randn(1, 8, 96, generator=g).to(dtype), torch.randn(1, 8, 96, generator=g).to(dtype)
        assert torch.equal(ops.silu_mul(gate, up), torch.nn.functional.silu(gate) * up)


def test_supports_and_gate(monkeypatch):
    m = _tiny(torch.float16)
    ids = torch.randint(0, 199, (1, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        h, kw = evalloop.layer_inputs(m, evalloop.LLAMA, ids)
    assert fused.supports(m, h, kw)
    assert not fused.supports(m, h.float(), kw)
    assert not fused.supports(m, torch.cat((h, h)), kw)
    monkeypatch.delenv('QUIP_FUSED_LAYER', raising=False)
    assert not fused.enabled()
    monkeypatch.setenv('QUIP_FUSED_LAYER', '1')
    assert fused.enabled()


def test_cuda_glue_refuses_cpu_tensors():
    x = torch.zeros(1, 4, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match='CUDA device only'):
        fused.CudaGlue().rmsnorm(x, torch.ones(64, dtype=torch.float16), 1e-5)


@pytest.mark.parametrize('nkv,use_ops', [(4, False), (2, False), (4, True), (2, True)])
def test_decoder_steps_equal_hf_decode_with_kv_cache_on_cpu(nkv, use_ops):
    """GraphDecoder's restated decode step (torch glue, and the fused-glue variant with the torch restatement injected)
    against the HF forward with a KV cache, eagerly on the CPU in fp32."""
    from quip_b200.decode import GraphDecoder
    m = _tiny(torch.float32, nkv)
    ids = torch.randint(0, 199, (2, 9), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        dec = GraphDecoder(m, max_len=16, batch=2, ops=TorchGlue() if use_ops else None)
        dec.cos, dec.sin = dec.cos.float(), dec.sin.float()                 # the decoder keeps fp16 tables for the GPU path
        with torch.no_grad():
            cos, sin = m.model.rotary_emb(torch.zeros(1, 1, 128), torch.arange(16)[None, :])
        dec.cos, dec.sin = cos[0].contiguous(), sin[0].contiguous()
        dec.k_cache, dec.v_cache = dec.k_cache.float(), dec.v_cache.float()
        past = None
        for i in range(ids.shape[1]):
            out = m(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            past = out.past_key_values
            got = dec.step(ids[:, i])
            assert torch.allclose(got, out.logits[:, -1], rtol=2e-4, atol=2e-4), i


def test_bench_glue_selection_falls_back_to_hf_and_respects_the_env(monkeypatch):
    """bench.pick_glue: an explicit QUIP_FUSED_LAYER wins; otherwise the fused stack must run, agree and be faster --
    on a CPU model the CUDA-only kernels raise, which keeps the HF glue (and says why)."""
    import bench
    m = _tiny(torch.float16)
    ids = torch.randint(0, 199, (1, 32), generator=torch.Generator().manual_seed(1))
    monkeypatch.setenv('QUIP_FUSED_LAYER', '1')
    assert bench.pick_glue(m, ids) == dict(mode='fused', chosen_by='QUIP_FUSED_LAYER=1')
    monkeypatch.setenv('QUIP_FUSED_LAYER', '0')
    assert bench.pick_glue(m, ids)['mode'] == 'hf'
    monkeypatch.delenv('QUIP_FUSED_LAYER')
    info = bench.pick_glue(m, ids)
    assert info['mode'] == 'hf' and 'CUDA device only' in info['why'] and os.environ['QUIP_FUSED_LAYER'] == '0'
    monkeypatch.delenv('QUIP_FUSED_LAYER')


def test_bench_glue_selection_happy_path_with_injected_kernels(monkeypatch):
    """The decision logic of bench.pick_glue with the torch restatement standing in for the CUDA kernels and a fake
    device timer: fused is chosen only when it agrees with the HF layers AND is faster."""
    import bench
    from quip_b200 import fused
    m = _tiny(torch.float16)
    ids = torch.randint(0, 199, (1, 32), generator=torch.Generator().manual_seed(1))
    monkeypatch.setattr(fused, 'CudaGlue', TorchGlue)
    clock = {'calls': 0}

    def fake_timer(fn, reps, warm):
        fn()
        clock['calls'] += 1
        return clock['ms'][clock['calls'] - 1]

    monkeypatch.setattr(bench, '_device_ms', fake_timer)
    monkeypatch.delenv('QUIP_FUSED_LAYER', raising=False)
    clock.update(calls=0, ms=[4.0, 2.0])                                  # hf, fused: fused faster
    info = bench.pick_glue(m, ids)
    assert info['mode'] == 'fused' and info['rel_err_vs_hf_layers'] < 1e-3 and os.environ['QUIP_FUSED_LAYER'] == '1'
    assert info['ms_per_layer_hf'] == 2.0 and info['ms_per_layer_fused'] == 1.0          # per layer: two layers timed
    monkeypatch.delenv('QUIP_FUSED_LAYER')
    clock.update(calls=0, ms=[2.0, 4.0])                                  # fused slower: keep the HF glue
    assert bench.pick_glue(m, ids)['mode'] == 'hf' and os.environ['QUIP_FUSED_LAYER'] == '0'
    monkeypatch.delenv('QUIP_FUSED_LAYER')

    class WrongGlue(TorchGlue):                                            # a kernel that disagrees: rejected however fast
        def silu_mul(self, gate, up):
            return gate * up

    monkeypatch.setattr(fused, 'CudaGlue', WrongGlue)
    clock.update(calls=0, ms=[4.0, 1.0])
    info = bench.pick_glue(m, ids)
    assert info['mode'] == 'hf' and info['rel_err_vs_hf_layers'] > 1e-3
    monkeypatch.delenv('QUIP_FUSED_LAYER')


def test_bench_decode_glue_check_logic(monkeypatch):
    import bench
    from quip_b200 import fused
    m = _tiny(torch.float32)
    m.half = lambda: m                                                   # keep fp32: CPU half matmuls are slow and noisy
    monkeypatch.setattr(fused, 'CudaGlue', TorchGlue)
    import quip_b200.decode as dec
    orig = dec.GraphDecoder.__init__

    def init32(self, *a, **kw):                                          # the decoder keeps fp16 tables for the GPU path
        orig(self, *a, **kw)
        cos, sin = m.model.rotary_emb(torch.zeros(1, 1, 128), torch.arange(self.max_len)[None, :])
        self.cos, self.sin = cos[0].contiguous(), sin[0].contiguous()
        self.k_cache, self.v_cache = self.k_cache.float(), self.v_cache.float()

    monkeypatch.setattr(dec.GraphDecoder, '__init__', init32)
    ok, worst, control = bench.decode_glue_ok(m, torch.device('cpu'))
    assert ok and worst < 1e-4

    class WrongGlue(TorchGlue):
        def silu_mul(self, gate, up):
            return gate * up

    monkeypatch.setattr(fused, 'CudaGlue', WrongGlue)
    ok, worst, control = bench.decode_glue_ok(m, torch.device('cpu'))
    assert not ok and worst > 2e-3 and worst > 3 * control
    monkeypatch.delenv('QUIP_FUSED_LAYER', raising=False)


@pytest.mark.parametrize('before', [True, False])
def test_opt_decoder_steps_equal_hf_decode_with_kv_cache_on_cpu(before):
    """GraphDecoder's OPT step (learned positions with the offset of 2, pre- / post-LayerNorm, scaled q, ReLU MLP, biases)
    against the HF OPT forward with a KV cache -- the loop the reference's benchmark() runs (opt.py:431-482) -- in fp32 on
    the CPU.  word_embed_proj_dim != hidden_size exercises project_in / project_out (OPT-350m's shape)."""
    from transformers import OPTConfig, OPTForCausalLM
    from quip_b200.decode import GraphDecoder
    torch.manual_seed(0)
    cfg = OPTConfig(hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=199,
                    max_position_embeddings=32, word_embed_proj_dim=64 if before else 48, do_layer_norm_before=before)
    m = OPTForCausalLM(cfg).float().eval()
    ids = torch.randint(0, 199, (2, 9), generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        dec = GraphDecoder(m, max_len=16, batch=2)
        dec.k_cache, dec.v_cache = dec.k_cache.float(), dec.v_cache.float()
        past = None
        for i in range(ids.shape[1]):
            out = m(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            past = out.past_key_values
            got = dec.step(ids[:, i])
            assert torch.allclose(got, out.logits[:, -1], rtol=2e-4, atol=2e-4), i
    with pytest.raises(AssertionError):
        GraphDecoder(m, max_len=64)                        # beyond the learned position table
