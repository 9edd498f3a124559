"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's block-render path (``np_oracle``), the harness
that imports the unmodified reference in the build container (``ref_harness``), the
shared parity cases (``cases``) and the golden-vector generator (``make_golden``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import anything from here; the product path (``signals_b200``) must not.
"""
