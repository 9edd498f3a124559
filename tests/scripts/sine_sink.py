"""Test entry point written for this repo in the style of the reference's scripts/edited_sine.py
(Rack -> SinkDevice -> Sine <- Fixed, then sink.start()), with the state-setting idiom of
scripts/edited_plot.py:23-24 so the oscillator is actually audible, plus a Gain so the graph equals
scripts/example_sine.py's formula (f = 500 Hz, a = 0.2).  Run through `python -m signals_b200.run_script`."""
import time

import numpy as np

import signals.chain.dev
import signals.chain.discovery
import signals.chain.fixed
import signals.chain.fx
import signals.chain.osc


def main():
    rack = signals.chain.discovery.Rack()
    rack.scan()
    choice = input('Enter device name: ')
    sink = signals.chain.dev.SinkDevice(rack.get_sink(choice))

    hertz = signals.chain.fixed.Fixed()
    hertz.get_state().value = np.array([500.0], ndmin=2)
    sine = signals.chain.osc.Sine()
    sine.hertz = hertz
    amp = signals.chain.fixed.Fixed()
    amp.get_state().value = np.array([0.2], ndmin=2)
    gain = signals.chain.fx.Gain()
    gain.left = sine
    gain.right = amp
    sink.input = gain

    sink.start()
    while sink.is_active:
        time.sleep(1)
    sink.destroy()


if __name__ == '__main__':
    main()
