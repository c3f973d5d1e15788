"""theta <-> T conversion, mirror of the reference `temperature` module (temperature.py:7-19)."""
import torch

from . import _host, _lib


def _convert(direction, t, p):
    fam = _host.Family(t, p)
    tt, pp = _host.dev(t), _host.dev(p)
    if tt.shape != pp.shape:                     # the reference asserts equal shapes (temperature.py:9, :16)
        tt, pp = (x.contiguous() for x in torch.broadcast_tensors(tt, pp))
    out = _host.empty(tt.shape)
    _lib.check(_lib.lib().gcm_temperature_convert(direction, _host.ptr(tt), _host.ptr(pp), _host.ptr(out), tt.numel(),
                                                  _lib.stream()), "gcm_temperature_convert")
    return fam.out(out, "kelvin")


def to_true_temp(t, p):
    """T = theta / (P0 / p)^kappa   (temperature.py:7-12)."""
    return _convert(0, t, p)


def to_potential_temp(t, p):
    """theta = T * (P0 / p)^kappa   (temperature.py:15-19)."""
    return _convert(1, t, p)
