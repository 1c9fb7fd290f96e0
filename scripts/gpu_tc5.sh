#!/bin/bash
timeout 120 python scripts/dbg_tc.py 2>&1 | tail -16
