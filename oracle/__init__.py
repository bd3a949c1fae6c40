"""CPU oracle for the Neural-Shadow-Mapping U-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pcss-unet_b200/`` (the product) may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

The reference (SDU-Gary/PCSS-Unet) is pure Python whose arithmetic lives in a third-party
dependency, PyTorch (pinned ``torch==2.5.1+cu124``, ``requirements.txt:9``; the image carries
2.11.0, op semantics unchanged).  The oracle therefore restates the reference's *call sequence*
functionally on top of ``torch.nn.functional`` CPU ops, with explicit parameter dictionaries,
each function citing the reference file:line it follows.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference classes
from ``/root/reference`` (build container only), runs them on seeded inputs and stores the outputs
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against those
vectors on every run (CPU, no GPU, no reference tree needed).
"""
from .unet_oracle import *  # noqa: F401,F403
from .stats_oracle import *  # noqa: F401,F403
