from .adan_sf import adan_sf  # noqa: F401
from .adamw import AdamW  # noqa: F401
from .fsam import fsam  # noqa: F401
