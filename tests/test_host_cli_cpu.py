"""CPU: the host side of `bkx-align` (option parsing, read parsing / packing, post-alignment passes, summary block,
CSV / BED / SAM / BAM writers) against the reference's own output files, with the C ABI answered by a test double built
on the oracle (tests/bkx_cpu_double.cpp).  The test bodies are the ones of tests/test_gpu_cli.py, which run the real
binary on the GPU box; only the executable differs.  Test infrastructure: the double is never shipped."""
import os
import subprocess

import pytest

import pyoracle as po
import test_gpu_cli as g
import test_gpu_zpefilter as gz

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="session")
def cpu_cli(tmp_path_factory):
    po.build()
    odir = os.path.join(ROOT, "oracle", "_build")
    exe = tmp_path_factory.mktemp("cpucli") / "bkx-align-cpu"
    csrc = os.path.join(ROOT, "biokanga_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(csrc, "host", "bkx_align_main.cpp"),
                    os.path.join(HERE, "bkx_cpu_double.cpp"), "-x", "c++", os.path.join(csrc, "bkx_cluster.cu"), "-x", "none",
                    "-L" + odir, "-lbkoracle", "-lz", "-lpthread", "-Wl,-rpath," + odir], check=True)
    return str(exe)


@pytest.fixture
def cli(cpu_cli, monkeypatch):
    monkeypatch.setattr(g, "CLI", cpu_cli)
    return cpu_cli


@pytest.mark.parametrize("case,tag", g.RUNS)
def test_host_outputs_match_reference(case, tag, cli, golden_dir, tmp_path):
    g.test_cli_outputs_match_reference(case, tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", gz.PE_FILTER_RUNS)
def test_host_paired_end_chromosome_filters_match_reference(tag, cli, golden_dir, tmp_path):
    gz.test_cli_paired_end_chromosome_filters_match_reference(tag, golden_dir, tmp_path)


def test_host_rejects_unsupported_and_bad_options(cli, golden_dir, tmp_path):
    g.test_cli_rejects_unsupported_and_bad_options(tmp_path, golden_dir)


@pytest.mark.parametrize("tag", ["m1", "m2", "m3", "m4", "m4t", "g0", "g1", "g2", "gz", "trim"])
def test_host_output_formats_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_output_formats_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.POST_TAGS)
def test_host_post_alignment_passes_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_post_alignment_passes_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["k0", "k50", "k250", "k20sam", "k0x4"])
def test_host_pcr_artefact_reduction_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_pcr_artefact_reduction_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["c5", "c5sam", "c5k", "c5pe", "c5pesam"])
def test_host_loci_base_constraints_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["s7", "s3pe", "pex0", "pex6", "p6", "p6sam", "p6pe", "p6pesam", "p6u1", "p6u3sam"])
def test_host_read_sampling_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_read_sampling_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["o1", "o2", "o3pe"])
def test_host_statistics_file_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_statistics_file_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["i1", "i2", "i3", "i4", "i5"])
def test_host_option_interplay_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_option_interplay_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.SIM_TAGS)
def test_host_simulated_read_truth_check_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_simulated_read_truth_check_matches_reference(tag, golden_dir, tmp_path)


def test_host_parameter_file_long_options_and_wildcards_match_reference(cli, golden_dir, tmp_path):
    g.test_cli_parameter_file_long_options_and_wildcards_match_reference(golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.BEST_TAGS)
def test_host_best_matches_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_best_matches_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.PRIORITY_TAGS)
def test_host_priority_regions_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_priority_regions_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.CONTAM_TAGS)
def test_host_adaptor_trimming_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_adaptor_trimming_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["e_m0", "e_m6", "e_m4x", "e_pe", "e_r5", "e_r4", "e_r1"])
def test_host_run_without_a_read_matches_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_run_without_a_read_matches_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", g.MANY_TAGS)
def test_host_hundreds_of_loci_per_read_match_reference(tag, cli, golden_dir, tmp_path):
    g.test_cli_hundreds_of_loci_per_read_match_reference(tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag,args,out", g.BAM_RUNS)
def test_host_bam_and_bai_match_reference(tag, args, out, cli, golden_dir, tmp_path):
    g.test_cli_bam_and_bai_match_reference(tag, args, out, golden_dir, tmp_path)


@pytest.mark.parametrize("case,tag", [("tiny", "mixed_s3"), ("tiny", "pe_U1")])
def test_host_chunked_parallel_parse_gives_the_same_files(case, tag, cli, golden_dir, tmp_path):
    g.test_cli_chunked_parallel_parse_gives_the_same_files(case, tag, golden_dir, tmp_path)


@pytest.mark.parametrize("tag", ["r5_R5_s3", "r5_R3_X_s3", "r5_R8_s5_e2"])
def test_host_all_loci_mode_sam_log_and_numbering(tag, cli, golden_dir, tmp_path):
    g.test_cli_all_loci_mode_sam_log_and_numbering(tag, golden_dir, tmp_path)
