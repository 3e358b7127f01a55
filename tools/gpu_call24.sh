#!/bin/bash
timeout 600 python tools/loss_profile.py 8 device 2>&1 | grep -v Warning | head -70
