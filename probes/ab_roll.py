"""A/B of two builds of the library on ONE box: usage ab_roll.py <lib.so> [layer ...] (runs probes/conv_layers.py with
b200seg._lib.LIB_PATH pointed at the given build)."""
import os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200seg._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [sys.argv[0]] + sys.argv[2:]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "conv_layers.py"), run_name="__main__")
