#!/bin/bash
# diagnostic: identify the box, run the PE fuzz, and on failure list the lost loci
nvidia-smi --query-gpu=name,serial,uuid,driver_version,vbios_version,ecc.errors.corrected.volatile.total,ecc.errors.uncorrected.volatile.total,clocks.max.sm --format=csv,noheader
hostname; uname -r
BKX_FUZZ_PE_SEEDS=40 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -k paired 2>&1 | tail -1
python tests/diag_lost_loci.py 2>&1 | tail -30 | cut -c1-330
