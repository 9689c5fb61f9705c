#!/bin/bash
timeout 500 python scripts/config_sweep.py 2>&1 | tail -8
