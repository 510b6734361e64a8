"""Imports the product package (its directory name has a hyphen, so it cannot be a plain `import`)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "markov-huffman-coding_b200"


def load():
    return importlib.import_module(PKG)
