#!/bin/bash
# Diagnostic for the open issue in DESIGN.md section 4: identify the box, run the paired-end fuzz, and if anything
# diverges list the loci the GPU did not see.  Index opens report self-check failures / retries on stderr.
nvidia-smi --query-gpu=name,serial,driver_version,vbios_version,ecc.errors.corrected.volatile.total,ecc.errors.uncorrected.volatile.total --format=csv,noheader
BKX_FUZZ_PE_SEEDS=${1:-40} python -m pytest tests/test_gpu_fuzz.py -m gpu -q -k paired 2>&1 | grep -E "DIAG|warning|passed|failed" | cut -c1-300 | tail -8
python tests/diag_lost_loci.py 2>&1 | grep -E "warning|seed|read|TOTAL" | tail -30 | cut -c1-330
