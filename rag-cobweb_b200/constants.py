"""Module-level switches with the reference's names (src/utils/constants.py)."""

# CobwebTorchTree.ifit takes "new" at every internal node instead of scoring the operations
# (src/cobweb/CobwebTorchTree.py:209-213, CobwebTorchNode.py:411-414).  Read when a tree is constructed.
COBWEB_GREEDY_MODE = False
