"""The reference-named entry points that had no test: opt_multigpu / llama_multigpu (opt.py:384-428, llama.py:361-415),
opt_pack3 / load_quant3 (opt.py:303-348), make_quant4 (zeroShot/models/quant.py:215-228), plus module copies / dtype casts.
CPU only (the packed forward itself needs a GPU: tests/test_gpu_*.py)."""
import copy
import io
import pickle

import pytest
import torch
import torch.nn as nn

from quip_b200 import quant as Q


def _tiny_opt(layers=4):
    from transformers import OPTConfig
    from quip_b200.opt import get_opt
    cfg = OPTConfig(vocab_size=128, hidden_size=128, ffn_dim=256, num_hidden_layers=layers, num_attention_heads=4,
                    max_position_embeddings=64, word_embed_proj_dim=128, do_layer_norm_before=True)
    torch.manual_seed(0)
    return get_opt(cfg, dtype=torch.float32)


def _tiny_llama(layers=4):
    from transformers import LlamaConfig
    from quip_b200.llama import get_llama
    cfg = LlamaConfig(vocab_size=128, hidden_size=128, intermediate_size=256, num_hidden_layers=layers, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64)
    torch.manual_seed(0)
    return get_llama(cfg, dtype=torch.float32, seqlen=32)


@pytest.mark.parametrize('family', ['opt', 'llama'])
def test_multigpu_placement_keeps_the_model_output(family):
    """MoveModule placement with the reference's contiguous ceil(L/G) rule; 'gpus' are CPU devices here, which exercises
    the wrapper (tensor and tuple kwargs moved to the stage's device) without hardware."""
    if family == 'opt':
        from quip_b200.opt import layer_placement, opt_multigpu as place
        m = _tiny_opt()
        layers = lambda mm: mm.model.decoder.layers        # noqa: E731
    else:
        from quip_b200.llama import llama_multigpu as place
        from quip_b200.opt import layer_placement
        m = _tiny_llama()
        layers = lambda mm: mm.model.layers                # noqa: E731
    ids = torch.randint(0, 128, (1, 16), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(ids).logits
    gpus = [torch.device('cpu')] * 3
    place(m, gpus)
    assert m.gpus == gpus
    wrapped = layers(m)
    assert all(type(w).__name__ == 'MoveModule' for w in wrapped)
    assert layer_placement(len(wrapped), 3) == [(0, 2), (2, 4), (4, 4)]          # ceil(4/3) = 2 per GPU, opt.py:424-426
    with torch.no_grad():
        got = m(ids).logits
    assert torch.allclose(got, want, atol=1e-5)


def test_llama_multigpu_layers_dist():
    from quip_b200.llama import llama_multigpu, parse_layers_dist
    m = _tiny_llama()
    llama_multigpu(m, [torch.device('cpu')] * 2, layers_dist='1:3')               # --layers-dist, llama.py:400-413
    assert parse_layers_dist('1:3', 4) == [(0, 1), (1, 4)]
    ids = torch.randint(0, 128, (1, 8), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert torch.isfinite(m(ids).logits).all()


def _grid_quantize_(model, names, bits):
    """Put every named Linear's weight on a per-channel asymmetric grid (what the reference's quantizer leaves behind)."""
    quantizers = {}
    mods = dict(model.named_modules())
    for n in names:
        lin = mods[n]
        qz = Q.Quantizer()
        qz.configure(bits, perchannel=True, sym=False, mse=False)
        qz.find_params(lin.weight.data, weight=True)
        lin.weight.data = Q.quantize_qfna(lin.weight.data, qz.scale, qz.zero, qz.maxq)
        quantizers[n] = qz
    return quantizers


def test_opt_pack3_and_load_quant3_round_trip():
    """opt_pack3(model, quantizers) (opt.py:303-315) swaps and packs every quantized Linear; its state_dict loads back
    through load_quant3 (opt.py:317-348) into packed modules with bit-identical codes."""
    from quip_b200.modelutils import find_layers
    from quip_b200.opt import load_quant3, opt_pack3
    m = _tiny_opt(layers=2)
    names = [n for n in find_layers(m) if 'layers' in n]
    quantizers = _grid_quantize_(m, names, 3)
    dense = {n: mod.weight.data.clone() for n, mod in find_layers(m).items() if n in names}
    opt_pack3(m, quantizers)
    packed = find_layers(m, [Q.QuantLinear])
    assert sorted(packed) == sorted(names) and all(p.bits == 3 for p in packed.values())
    for n in names[:3]:
        ql = packed[n]
        w = ql.scales * ql.codes().float() - ql.zeros                                # Q = scales*codes - zeros
        assert torch.allclose(w, dense[n], atol=1e-5)
    buf = io.BytesIO()
    torch.save(m.state_dict(), buf)
    buf.seek(0)
    m2 = load_quant3(m.config, torch.load(buf))
    packed2 = find_layers(m2, [Q.QuantLinear])
    assert sorted(packed2) == sorted(names)
    for n in names:
        assert torch.equal(packed2[n].qweight, packed[n].qweight)


def test_make_quant4_packs_from_the_quantizers():
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.head = nn.Linear(128, 256), nn.Linear(256, 128), nn.Linear(128, 32)

    torch.manual_seed(0)
    net = Net()
    for lin in (net.fc1, net.fc2, net.head):          # load_quant (like the reference's) turns torch's initialisers into no-ops
        lin.weight.data = torch.randn_like(lin.weight) * 0.05
        lin.bias.data = torch.randn_like(lin.bias) * 0.1
    quantizers = _grid_quantize_(net, ['fc1', 'fc2'], 4)
    dense = {n: getattr(net, n).weight.data.clone() for n in ('fc1', 'fc2')}
    Q.make_quant4(net, quantizers)                                                   # zeroShot/models/quant.py:215-228
    assert isinstance(net.fc1, Q.QuantLinear) and net.fc1.bits == 4 and type(net.head) is nn.Linear
    for n in ('fc1', 'fc2'):
        ql = getattr(net, n)
        assert torch.allclose(ql.scales * ql.codes().float() - ql.zeros, dense[n], atol=1e-5)


def test_packed_module_survives_dtype_casts_copies_and_pickling():
    """model.half() / .float() must not change the dtypes the C ABI reads (scales fp32, factors fp16, ...), and a module
    that has built its ctypes descriptor can still be deep-copied, pickled and torch.save'd."""
    from conftest import load_layer, parts_to_torch
    parts, _ = load_layer('l2b_incoh')
    tp = parts_to_torch(parts)
    N, K = tp.codes.shape
    ql = Q.QuantLinear(infeatures=K, outfeatures=N, **Q.spec_from_parts(tp))
    ql.pack_parts(tp)
    want = {n: b.dtype for n, b in ql.named_buffers()}
    for cast in (lambda m: m.half(), lambda m: m.float(), lambda m: m.to(torch.bfloat16), lambda m: m.double()):
        cast(ql)
        assert {n: b.dtype for n, b in ql.named_buffers()} == want
    ql._descriptor()                                                                 # ctypes pointers are now cached
    assert ql._desc is not None
    for clone in (copy.deepcopy(ql), pickle.loads(pickle.dumps(ql))):
        assert clone._desc is None and clone._group is None
        assert torch.equal(clone.qweight, ql.qweight) and clone.qweight.data_ptr() != ql.qweight.data_ptr()
        clone._descriptor()
    buf = io.BytesIO()
    torch.save(ql, buf)
    ql.scales = ql.scales.half()                                                     # a wrong dtype is refused, not reinterpreted
    ql._desc = None
    with pytest.raises(TypeError, match='scales'):
        ql._descriptor()
