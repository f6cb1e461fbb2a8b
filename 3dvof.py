#!/usr/bin/env python
"""`python 3dvof.py [-ic {1,2,3}] [-s]` -- same command line as the reference's 3-D script, executed by the B200-native
library (see taichi_2d_vof_b200/driver3d.py for the extensions)."""
import sys

from taichi_2d_vof_b200.driver3d import main

if __name__ == "__main__":
    sys.exit(main())
