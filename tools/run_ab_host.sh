#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/hostbreak.py 2>&1 | tail -16
tools/ab.sh "c2 c1 c3"
