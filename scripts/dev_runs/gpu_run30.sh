#!/bin/bash
QR_TRACE=1 timeout 200 python scripts/longrun.py 203 2>&1 | grep -E "trace|trees" | tail -8
QR_TRACE=1 QR_TRACE_ROUNDS=1 timeout 200 python scripts/longrun.py 201 2>&1 | grep -E "trace" | tail -45
