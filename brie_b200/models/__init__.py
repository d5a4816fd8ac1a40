"""Public re-exports, as brie/models/__init__.py:1-4."""
from .model_b200 import BRIE2, Model_init
from .model_wrap import fit_BRIE_matrix, fitBRIE, BRIE_RV, concate
from .base_model import get_CI95, BRIE_base_lik, LogitNormal
