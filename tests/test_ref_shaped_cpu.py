"""The reference-shaped stand-in stages (tests/ref_shaped.py) are pinned here against the REAL reference stage's fixture: with the
reference's state_dict loaded and the oracle solver behind the module-global ``sdeint`` they reproduce SDEDecoder.forward's outputs."""
import torch

import ref_shaped
from oracle import sde_oracle as so


def _oracle_sdeint(dW):
    def f(sde, y0, ts, **kw):
        assert kw['method'] == 'euler' and kw['dt'] == 0.1
        pf = {k: v for k, v in sde.f_func.net.state_dict().items()}
        pg = {k: v for k, v in sde.g_func.net.state_dict().items()}
        return so.euler_solve_ref(pf, pg, y0, ts, kw['dt'], dW)[0]
    return f


def test_ref_shaped_decoder_reproduces_the_reference_stage_fixture(golden_stage, monkeypatch):
    d = golden_stage
    sd = {k[len('param/'):]: torch.from_numpy(d[k]) for k in d if k.startswith('param/')}
    dec = ref_shaped.RefShapedDecoder().load_reference_state_dict(sd)
    assert set(dec.state_dict()) == {k for k in sd if not k.startswith('lsde_func.h_func')}
    monkeypatch.setattr(ref_shaped, 'sdeint', _oracle_sdeint(torch.from_numpy(d['dW'])))
    with torch.no_grad():
        out = dec({'padding_mask': torch.from_numpy(d['padding_mask'])}, torch.from_numpy(d['local_embed']), torch.from_numpy(d['global_embed']))
    assert torch.allclose(out['loc'], torch.from_numpy(d['loc']), atol=2e-5, rtol=1e-5)
    assert torch.allclose(out['pi'], torch.from_numpy(d['pi']), atol=1e-5, rtol=1e-5)
    assert torch.equal(out['reg_mask'], torch.from_numpy(d['reg_mask']))


def test_ref_shaped_encoder_state_dict_matches_reference_names():
    """Parameter names of the SDE / GRU part equal the real LocalEncoderSDESepPara2's (`gru_unit.*`, `lsde_func.f_func.net.*`,
    `lsde_func.g_nus.net.*`, `lsde_func.g_argo.net.*`, `hidden`); checked against the live class when the reference tree is present."""
    enc = ref_shaped.RefShapedEncoder()
    keys = {k for k in enc.state_dict() if k.startswith(('gru_unit.', 'lsde_func.', 'hidden'))}
    assert 'gru_unit.update_gate.0.weight' in keys and 'lsde_func.g_argo.net.4.bias' in keys and 'hidden' in keys
    from oracle import ref_runner as rr
    if rr.reference_available():
        from test_boundary_cpu import REF_ENC_KW
        ref = rr.load_reference()['enc'].LocalEncoderSDESepPara2(**REF_ENC_KW)
        ref_keys = {k for k in ref.state_dict() if k.startswith(('gru_unit.', 'lsde_func.', 'hidden')) and 'h_func' not in k}
        assert keys == ref_keys
