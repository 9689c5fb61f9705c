#!/bin/bash
echo "host-driven"; timeout 200 python scripts/longrun.py 400 2>&1 | tail -8
echo "device-driven"; QR_DEVICE_GROWTH=1 timeout 200 python scripts/longrun.py 400 2>&1 | tail -8
