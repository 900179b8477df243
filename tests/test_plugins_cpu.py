"""The YAML-selectable stage files (trajsde_b200/plugins) load through the reference's own loader (model_base_mix_sde.py:38-45) and
keep the reference classes' constructor and state_dict (dev container only: needs /root/reference)."""
import os
from importlib.machinery import SourceFileLoader

import pytest

from oracle import ref_runner as rr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(module_name, rel):
    return getattr(SourceFileLoader(module_name, os.path.join(ROOT, rel)).load_module(module_name), module_name)


@pytest.mark.skipif(not rr.reference_available(), reason="reference tree absent (GPU box)")
def test_plugin_stage_classes_keep_the_reference_state_dict():
    from test_boundary_cpu import REF_ENC_KW
    from trajsde_b200.stages import FusedDecoderMixin, FusedEncoderMixin
    m = rr.load_reference()
    Dec = _load('SDEDecoderFused', 'trajsde_b200/plugins/dec_hivt_nusargo_sde_fused.py')
    dec, ref = Dec(**rr.DEC_KW), m['dec'].SDEDecoder(**rr.DEC_KW)
    from models.decoders.dec_hivt_nusargo_sde import SDEDecoder as RefDecoder        # the class the plugin file subclasses
    assert isinstance(dec, RefDecoder) and type(dec).forward is FusedDecoderMixin.forward
    assert list(dec.state_dict()) == list(ref.state_dict())
    dec.load_state_dict(ref.state_dict())
    Enc = _load('LocalEncoderSDESepPara2Fused', 'trajsde_b200/plugins/enc_hivt_nusargo_sde_sep2_fused.py')
    enc, renc = Enc(**REF_ENC_KW), m['enc'].LocalEncoderSDESepPara2(**REF_ENC_KW)
    assert type(enc).forward is FusedEncoderMixin.forward and type(enc).forward_ood is FusedEncoderMixin.forward_ood
    assert list(enc.state_dict()) == list(renc.state_dict())
    enc.load_state_dict(renc.state_dict())


@pytest.mark.skipif(not rr.reference_available(), reason="reference tree absent (GPU box)")
def test_plugin_encoder_prepare_and_finish_equal_the_reference_forward(golden_encoder_stage):
    """The PyG-dependent code of the YAML-selectable encoder class (`_prepare`: perturbed agent copies, per-slot subgraphs, AA encoder;
    `_finish`: AL encoder) against what the reference's own forward computed at the same points (tests/golden/encoder_stage.npz was
    captured from it): same weights, same global-RNG seed for the perturbed copies, functional PyG stand-in of oracle/shims."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden as mg
    d = golden_encoder_stage
    Enc = _load('LocalEncoderSDESepPara2Fused', 'trajsde_b200/plugins/enc_hivt_nusargo_sde_sep2_fused.py')
    torch.manual_seed(int(d['init_seed']))
    enc = Enc(**mg.ENC_KW).eval()                               # same construction order as the fixture generator -> same HiVT weights
    rr._perturb_biases(enc.gru_unit, 0.1, torch.Generator().manual_seed(22))
    rr._perturb_biases(enc.lsde_func, 0.1, torch.Generator().manual_seed(23))
    for k in d:
        if k.startswith('param/'):
            assert torch.equal(enc.state_dict()[k[6:]], torch.from_numpy(d[k])), k
    data = mg.synthetic_graph_batch()
    torch.manual_seed(int(d['torch_seed_fake_agents']))
    with torch.no_grad():
        prep = enc._prepare(data, ood=False)
        assert torch.allclose(prep['aa_out'], torch.from_numpy(d['aa_out']), atol=1e-6, rtol=1e-5)
        assert torch.equal(prep['actors_mask'], torch.from_numpy(d['actors_mask'])) and torch.equal(prep['nus_mask'], torch.from_numpy(d['nus_mask']))
        assert prep['n_fake'] == 4 and torch.equal(prep['agent_index'], torch.from_numpy(d['agent_index']))
        out = enc._finish(data, prep, torch.from_numpy(d['pre_al']))
        assert torch.allclose(out, torch.from_numpy(d['out']), atol=1e-6, rtol=1e-5)
        prep_ood = enc._prepare(data, ood=True)                 # forward_ood: no perturbed copies, original graph
        assert prep_ood['aa_out'].shape == (21, 24, 64) and prep_ood['n_fake'] == 0
