"""CPU: a short run of the randomised front-end comparison (tests/fuzz_host_cli.py) -- the host side of bkx-align, through
the oracle-backed test double, against the reference binary itself on drawn option combinations.  Skipped where the
reference binary is not built (oracle/_ref, see oracle/build_ref.sh); longer runs: `python tests/fuzz_host_cli.py --seeds 0:300`."""
import os

import pytest

import fuzz_host_cli as fz


@pytest.mark.skipif(not os.path.exists(fz.REF), reason="reference binary not built (oracle/_ref)")
def test_front_end_matches_reference_on_drawn_option_sets(tmp_path):
    import gzip
    import shutil
    from concurrent.futures import ThreadPoolExecutor
    work = str(tmp_path)
    for f in ("tiny.sfx", "r100.fa", "r50.fa", "r150.fa", "mixed.fq", "pe1.fa", "pe2.fa"):
        with gzip.open(os.path.join(fz.GOLD, "tiny", f + ".gz"), "rb") as a, open(os.path.join(work, f), "wb") as b:
            shutil.copyfileobj(a, b)
    fz.write_side_files(work)
    cli = os.path.join(work, "bkx-align-cpu")
    fz.build_cpu_cli(cli)
    with ThreadPoolExecutor(8) as ex:
        bad = [r for r in ex.map(lambda s: fz.one(s, cli, work), range(1000, 1016)) if r]
    assert not bad, bad
