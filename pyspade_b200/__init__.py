"""spade-b200: B200-native engine for pySpade's hypergeometric DE hot path.

Drop-in surface (same names and argument order as ``pySpade/utils.py``):
``pyspade_b200.utils.perform_DE`` / ``perform_DE_NC`` / ``cpm_normalization``,
``pyspade_b200.DEobs.DE_observe_cells``, ``pyspade_b200.DErand.DE_random_cells``
and the ``DEobs`` / ``DErand`` sub-commands of ``python -m pyspade_b200``.
"""
__version__ = "0.1.0"
