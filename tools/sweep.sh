#!/bin/bash
cd "$(dirname "$0")/.."
python tools/prof_step.py --map squarinth --worlds 4096
python tools/prof_step.py --map squarinth --worlds 16384
python tools/prof_step.py --map agh-map --free 1 --worlds 16384
python tools/prof_step.py --map labyrinth --free 1 --worlds 8192
python tools/prof_step.py --map grandbyrinth --worlds 16384
