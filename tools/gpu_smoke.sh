#!/bin/bash
# the driver's round-end sequence in small: smoke() and the known-answer / golden parity tests
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q -x -k "kats or golden or sweep" 2>&1 | tail -2
