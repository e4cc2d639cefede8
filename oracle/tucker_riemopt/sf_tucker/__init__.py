from .sf_tucker import SFTucker  # noqa: F401
