"""Launch file with the reference's path (`python stoix/systems/ppo/anakin/rec_ppo.py k=v ...`); the implementation is
stoix_b200/systems/ppo/anakin/rec_ppo.py, and `import stoix.systems.ppo.anakin.rec_ppo` yields that very module."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from stoix_b200.systems.ppo.anakin import rec_ppo as _impl  # noqa: E402

if __name__ == "__main__":
    _impl.hydra_entry_point()
