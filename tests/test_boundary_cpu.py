"""Host-side contract of the drop-in boundary (no GPU): argument validation mirrors the reference's check_contract
(models/utils/sdeint.py:827-995); unsupported configurations raise instead of falling back."""
import types

import pytest
import torch

import trajsde_b200 as tb
from helpers import DecoderSDE, EncoderSDE
from trajsde_b200 import patch


class _FakeCuda(torch.Tensor):
    pass


def test_sdeint_contract_errors_cpu():
    sde = DecoderSDE()
    ts = torch.linspace(0, 1, 11)
    with pytest.raises(ValueError, match="must be a torch.Tensor"):
        tb.sdeint(sde, [[0.0] * 64], ts, dt=0.1, method='euler')
    with pytest.raises(ValueError, match="2-dimensional"):
        tb.sdeint(sde, torch.zeros(64), ts, dt=0.1, method='euler')
    with pytest.raises(NotImplementedError):
        tb.sdeint(sde, torch.zeros(2, 64), ts, dt=0.1, method='euler', adaptive=True)
    with pytest.raises(NotImplementedError):
        tb.sdeint(sde, torch.zeros(2, 64), ts, dt=0.1, method='euler', logqp=True)
    bad = DecoderSDE()
    bad.noise_type = 'general'
    with pytest.raises(NotImplementedError):
        tb.sdeint(bad, torch.zeros(2, 64), ts, dt=0.1, method='euler')
    del bad.noise_type
    type(bad).noise_type  # class attr still there; emulate missing attribute with a bare object
    with pytest.raises(ValueError, match="noise_type"):
        tb.sdeint(types.SimpleNamespace(sde_type='ito'), torch.zeros(2, 64), ts, dt=0.1, method='euler')


def test_default_mode_and_seed_api():
    assert tb.get_default_mode() in ('tc_f16', 'exact')
    old = tb.get_default_mode()
    tb.set_default_mode('exact')
    assert tb.get_default_mode() == 'exact'
    with pytest.raises(ValueError):
        tb.set_default_mode('fp8')
    tb.set_default_mode(old)
    tb.manual_seed(123)
    from trajsde_b200 import solver as mod
    a, b = mod._next_call_seed(), mod._next_call_seed()
    tb.manual_seed(123)
    assert (a, b) == (mod._next_call_seed(), mod._next_call_seed()) and a != b


def test_install_rebinds_reference_module_globals():
    """patch.install swaps the two module globals the reference stages call (dec…sde.py:11,88; enc…sep2.py:23,149)."""
    def make(name, glob_name):
        g = {glob_name: 'ORIGINAL', '__name__': name}
        src = f"class Stage:\n    def forward(self):\n        return {glob_name}\n"
        exec(src, g)
        return g['Stage'](), g

    dec, gd = make('SDEDecoder', 'sdeint')
    enc, ge = make('LocalEncoderSDESepPara2', 'sdeint_dual')
    model = types.SimpleNamespace(decoder=dec, encoder=enc)
    saved = patch.install(model)
    assert dec.forward() is patch._sdeint_rows_major and enc.forward() is tb.sdeint_dual     # decoder: sdeint with rows-major storage
    patch.uninstall(saved)
    assert dec.forward() == 'ORIGINAL' and enc.forward() == 'ORIGINAL'
    with pytest.raises(KeyError):
        patch.install(decoder=enc)


def test_install_binds_fused_gru_forward_on_the_instance_only():
    """install() also routes encoder.gru_unit.forward (the jump at enc…sep2.py:165-169) to the fused operator — as an instance
    attribute, so the reference class and other instances keep their forward; uninstall() removes it again."""
    from trajsde_b200 import synthetic as syn
    g = {'sdeint_dual': 'ORIGINAL', '__name__': 'LocalEncoderSDESepPara2'}
    exec("class Stage:\n    def forward(self):\n        return sdeint_dual\n", g)
    enc = g['Stage']()
    enc.gru_unit = syn.GRUUnit()
    other = syn.GRUUnit()
    saved = patch.install(encoder=enc)
    assert 'forward' in enc.gru_unit.__dict__ and 'forward' not in other.__dict__
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        enc.gru_unit(h_cur=torch.zeros(3, 64), input_tensor=torch.zeros(3, 64), mask=torch.ones(3, dtype=torch.bool))
    patch.uninstall(saved)
    assert 'forward' not in enc.gru_unit.__dict__
    assert enc.gru_unit(h_cur=torch.zeros(3, 64), input_tensor=torch.zeros(3, 64), mask=torch.ones(3, dtype=torch.bool)).shape == (3, 64)
    enc.gru_unit = syn.GRUUnit(n_units=100)                         # other widths stay on the reference path
    assert 'gru' not in patch.install(encoder=enc, fuse_gru=True)


def test_unsupported_net_layout_raises():
    from trajsde_b200.solver import _mlp_params
    import torch.nn as nn
    n = types.SimpleNamespace(net=nn.Sequential(nn.Linear(66, 64), nn.ReLU(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 64)))
    with pytest.raises(NotImplementedError):
        _mlp_params(n, 64, 'f')
    n = types.SimpleNamespace(net=nn.Sequential(nn.Linear(34, 32), nn.Tanh(), nn.Linear(32, 32), nn.Tanh(), nn.Linear(32, 32)))
    with pytest.raises(NotImplementedError):
        _mlp_params(n, 32, 'f')
    assert len(_mlp_params(EncoderSDE().g_argo, 1, 'g')) == 6


def test_install_binds_fused_heads_on_the_instances_only():
    """install() binds the fused head pair on decoder.decoder / decoder.scale (dec…sde.py:50-61) as instance attributes; inputs the
    fused launch does not serve (autograd, non-CUDA) run the reference's own nn.Sequential; uninstall() restores the class forward."""
    import torch.nn as nn

    def head():
        return nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2))

    g = {'sdeint': 'ORIGINAL', '__name__': 'SDEDecoder', 'nn': nn, 'head': head}
    exec("class Stage(nn.Module):\n    def __init__(self):\n        super().__init__()\n        self.decoder, self.scale = head(), head()\n"
         "    def forward(self):\n        return sdeint\n", g)
    dec = g['Stage']()
    x = torch.randn(3, 5, 64)
    want = dec.decoder(x)
    saved = patch.install(decoder=dec)
    assert 'heads' in saved and 'forward' in dec.decoder.__dict__ and 'forward' in dec.scale.__dict__
    assert torch.equal(dec.decoder(x), want)                  # CPU tensor: the reference module's own forward
    patch.uninstall(saved)
    assert 'forward' not in dec.decoder.__dict__ and 'forward' not in dec.scale.__dict__
    narrow = g['Stage']()
    narrow.decoder = nn.Sequential(nn.Linear(32, 32), nn.LayerNorm(32), nn.ReLU(), nn.Linear(32, 2))
    assert 'heads' not in patch.install(decoder=narrow)       # other widths: left alone


def test_install_on_the_live_reference_decoder():
    """With /root/reference present (dev container): install() on a real `SDEDecoder` instance swaps the module global its forward
    calls (dec…sde.py:11, 88) and binds the fused head pair on its `decoder` / `scale` instances; uninstall() restores both."""
    from oracle import ref_runner as rr
    if not rr.reference_available():
        pytest.skip("reference tree absent (GPU box)")
    dec = rr.build_reference_decoder(seed=0)
    g = type(dec).forward.__globals__
    orig = g['sdeint']
    saved = patch.install(decoder=dec)
    try:
        assert g['sdeint'] is patch._sdeint_rows_major and g['sdeint'] is not orig
        assert 'heads' in saved and 'forward' in dec.decoder.__dict__ and 'forward' in dec.scale.__dict__
        x = torch.randn(4, 60, 64)
        assert torch.equal(dec.decoder(x), torch.nn.Sequential.forward(dec.decoder, x))      # CPU input: the reference module itself
    finally:
        patch.uninstall(saved)
    assert g['sdeint'] is orig and 'forward' not in dec.decoder.__dict__ and 'forward' not in dec.scale.__dict__


REF_ENC_KW = dict(historical_steps=21, node_dim=2, edge_dim=2, embed_dim=64, num_heads=8, dropout=0.1, parallel=True, local_radius=50,
                  sde_layers=2, ref_time=20, max_past_t=2, run_backwards=True, minimum_step=0.1, rtol=0.001, atol=0.001, method='euler')


def test_install_on_the_live_reference_encoder_binds_the_gru_jump():
    """The reference encoder names its jump module `gru_unit` (enc…sep2.py:49; state_dict `encoder.gru_unit.*`): install() on a real
    `LocalEncoderSDESepPara2` instance must rebind `sdeint_dual` AND the GRU jump, and uninstall() must restore both."""
    from oracle import ref_runner as rr
    if not rr.reference_available():
        pytest.skip("reference tree absent (GPU box)")
    enc = rr.load_reference()['enc'].LocalEncoderSDESepPara2(**REF_ENC_KW)
    g = type(enc).forward.__globals__
    orig = g['sdeint_dual']
    saved = patch.install(encoder=enc)
    try:
        assert g['sdeint_dual'] is tb.sdeint_dual
        assert 'gru' in saved and 'forward' in enc.gru_unit.__dict__
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            enc.gru_unit(h_cur=torch.zeros(3, 64), input_tensor=torch.zeros(3, 64), mask=torch.ones(3, dtype=torch.bool))
    finally:
        patch.uninstall(saved)
    assert g['sdeint_dual'] is orig and 'forward' not in enc.gru_unit.__dict__
    assert any(k.startswith('gru_unit.') for k in enc.state_dict())


def test_default_seed_follows_torch_seed_and_rank(monkeypatch):
    """bm=None default seeding: derived from torch.initial_seed() (so torch.manual_seed / pl.seed_everything reproduce a run and an
    unseeded process draws fresh noise, like BrownianInterval(entropy=None)), mixed with the distributed rank."""
    from trajsde_b200 import solver as mod
    saved = dict(mod._defaults)
    try:
        mod._defaults.update(seed=None, torch_seed=None, calls=0)
        torch.manual_seed(1234)
        a = [mod._next_call_seed() for _ in range(3)]
        torch.manual_seed(1234)
        mod._defaults.update(seed=None, torch_seed=None, calls=0)
        assert a == [mod._next_call_seed() for _ in range(3)] and len(set(a)) == 3
        torch.manual_seed(99)                                    # re-seeding torch mid-run restarts the stream from the new seed
        b = mod._next_call_seed()
        assert b not in a
        torch.manual_seed(1234)
        assert mod._next_call_seed() == a[0]
        monkeypatch.setattr(mod, '_rank', lambda: 3)
        mod._defaults.update(seed=None, torch_seed=None, calls=0)
        assert mod._next_call_seed() != a[0]                     # another rank, another stream
        tb.manual_seed(5)                                        # explicit library seed: torch's seed no longer matters
        c = mod._next_call_seed()
        torch.manual_seed(7)
        tb.manual_seed(5)
        assert mod._next_call_seed() == c
    finally:
        mod._defaults.update(saved)


def test_device_schedule_cache_is_keyed_by_content():
    import inspect
    from trajsde_b200 import ops
    src = inspect.getsource(ops.DeviceSchedule.get)
    assert 'id(' not in src and 'tobytes' in src


def test_adjoint_range_policy_api():
    from trajsde_b200 import ops
    with pytest.raises(ValueError):
        ops.set_adjoint_range_policy('maybe')
    ops.set_adjoint_range_policy('raise')
    ops.set_adjoint_range_policy('warn')
    assert issubclass(ops.AdjointRangeError, FloatingPointError)
    ops.poll_status()                                             # no device touched yet: a no-op


def test_fused_head_pair_key_handles_inference_tensors():
    """Lightning runs validate/test/predict under torch.inference_mode(): inference tensors have no version counter."""
    from trajsde_b200.heads import FusedHeadPair
    with torch.inference_mode():
        x = torch.zeros(2, 3, 64)
        assert FusedHeadPair._key_of(x)[1] is None
    y = torch.zeros(2, 3, 64)
    assert FusedHeadPair._key_of(y)[1] == y._version
