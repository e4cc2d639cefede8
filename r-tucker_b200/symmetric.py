"""Drop-in for ``src/model/symmetric/{R_TuckER,optim}.py`` of the reference."""
from .model import SymmetricRTuckER as R_TuckER  # noqa: F401
from .optim import SymRGD as RGD, SymRSGDwithMomentum as RSGDwithMomentum  # noqa: F401
from .optim import SFTuckerAdam as RiemannianAdam, SFTuckerAdam  # noqa: F401
from .manifold import SFTucker  # noqa: F401
