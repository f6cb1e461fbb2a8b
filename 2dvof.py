#!/usr/bin/env python
"""`python 2dvof.py [-ic {1,2,3}] [-s]` -- same command line as the reference script, executed by
the B200-native library (see taichi_2d_vof_b200/driver.py for the extensions)."""
import sys

from taichi_2d_vof_b200.driver import main

if __name__ == "__main__":
    sys.exit(main())
