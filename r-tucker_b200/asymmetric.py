"""Drop-in for ``src/model/asymmetric/{R_TuckER,optim}.py`` of the reference."""
from .model import AsymmetricRTuckER as R_TuckER  # noqa: F401
from .optim import AsymRGD as RGD, AsymRSGDwithMomentum as RSGDwithMomentum  # noqa: F401
from .optim import TuckerAdam as RiemannianAdam, TuckerAdam  # noqa: F401
from .manifold import Tucker  # noqa: F401
