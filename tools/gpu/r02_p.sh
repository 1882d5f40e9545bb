#!/usr/bin/env bash
# Round 2, call P: slice shapes for one rank's share of a 2-, 4-, 8-way split and for the whole frame (world 1)
set -u
cd "$(dirname "$0")/../.."
for w in 2 1 4 8; do timeout 300 python tools/time_share.py $w 2>&1 | grep "^world"; done
