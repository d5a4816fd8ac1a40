"""brie_b200: B200-native BRIE2 variational-inference fit (drop-in for the hot
path of huangyh09/brie's `brie-quant`: brie.models.BRIE2 / fit_BRIE_matrix / fitBRIE).
"""
from .version import __version__
