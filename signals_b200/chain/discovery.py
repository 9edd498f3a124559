"""Plugin discovery (mirror of signals.chain.discovery, /root/reference/src/signals/chain/discovery.py).

``load_signal`` resolves the dotted class names used by the patch language (``+ 2a signals.chain.osc.Sine``,
map/control.py:291-330).  Names under ``signals.`` resolve to this package's mirrors, so patch files written
for the reference load unchanged; anything else is resolved as is (third-party node modules).
"""
from __future__ import annotations

import abc
import pkgutil

from signals_b200 import SignalsError
from signals_b200.chain import Signal
from signals_b200.chain.dev import DeviceInfo, _sd


class DiscoveryError(SignalsError):
    pass


class BadSyntax(DiscoveryError):
    pass


class BadPath(DiscoveryError):
    pass


class InvalidObject(DiscoveryError):
    pass


class BadDeviceName(DiscoveryError):
    pass


class NotASource(DiscoveryError):
    pass


class NotASink(DiscoveryError):
    pass


def mirror_name(qualname: str) -> str:
    """``signals.chain.osc.Sine`` -> ``signals_b200.chain.osc.Sine`` (the reference's own package name)."""
    if qualname == 'signals' or qualname.startswith('signals.'):
        return 'signals_b200' + qualname[len('signals'):]
    return qualname


def load_signal(qualname: str) -> type:
    try:
        cls = pkgutil.resolve_name(mirror_name(qualname))
    except ValueError:
        raise BadSyntax(qualname)
    except (AttributeError, ImportError) as e:
        raise BadPath(qualname, e.args[0] if e.args else '')
    # concrete = no abstract methods left and not itself declared as an ABC (the mirror's node bases share one
    # GPU `_eval`, so unlike the reference's they have no abstract method to tell them apart)
    if (isinstance(cls, type) and issubclass(cls, Signal) and not getattr(cls, '__abstractmethods__', None)
            and abc.ABC not in cls.__bases__):
        return cls
    raise InvalidObject(qualname, cls)


class Rack:
    """Audio devices by name (discovery.py:96-126)."""

    def __init__(self):
        self.devices: list[DeviceInfo] = []

    def scan(self) -> None:
        self.devices[:] = (DeviceInfo(**info) for info in _sd().query_devices())

    def get_device(self, name: str) -> DeviceInfo:
        found = [d for d in self.devices if d.name == name]
        if len(found) != 1:
            raise BadDeviceName(name)
        return found[0]

    def get_source(self, name: str) -> DeviceInfo:
        device = self.get_device(name)
        if not device.is_source:
            raise NotASource(name)
        return device

    def get_sink(self, name: str) -> DeviceInfo:
        device = self.get_device(name)
        if not device.is_sink:
            raise NotASink(name)
        return device

    def sources(self) -> list[DeviceInfo]:
        return sorted(d for d in self.devices if d.is_source)

    def sinks(self) -> list[DeviceInfo]:
        return sorted(d for d in self.devices if d.is_sink)
