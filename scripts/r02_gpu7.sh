#!/bin/bash
QR_KTRACE=20 timeout 300 python scripts/longrun.py 150 2>&1 | grep ktrace | tail -4
