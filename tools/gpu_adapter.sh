#!/bin/bash
timeout 900 python -m pytest tests/test_adapter.py -x -q 2>&1 | tail -8
