#!/bin/bash
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout -s KILL 300 python tools/hostbreak.py 2>&1 | tail -14
