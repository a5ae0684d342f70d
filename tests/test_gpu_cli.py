"""GPU: the host front-end `bkx-align` (drop-in for `biokanga align`) against the reference's own output
files: CSV rows, SAM header + records and the alignment-summary log block, for SE and PE runs."""
import gzip
import os
import subprocess

import pytest

import goldutil as gu
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu
CLI = os.path.join(os.path.dirname(bkx.LIB_PATH), "bkx-align")

RUNS = [("tiny", "r100_s3"), ("tiny", "r100_s5_e2"), ("tiny", "mixed_s3"), ("tiny", "r251_s6"), ("tiny", "r100_s3_Q2"),
        ("tiny", "r100_s4_m2"), ("tiny", "pe_U2"), ("tiny", "pe_U4"), ("tiny", "pe_U1"), ("tiny", "pe_U3"),
        ("tiny", "pe_U1_far"), ("repeats", "r100_s3_m3"), ("repeats", "r60_s5"), ("lowcopy", "r1_R5_s3"),
        ("lowcopy", "r1_R4_X_s3"), ("lowcopy", "r1_R20_s5_e2"), ("lowcopy", "r5_R5_s3"), ("lowcopy", "r5_R3_X_s3"),
        ("lowcopy", "r5_R8_s5_e2"), ("lowcopy", "r3_R5_s3"), ("lowcopy", "r4_R5_s3"), ("lowcopy", "r4_R8_X_s5")]


def summary_block(path):
    keep, on = [], False
    for ln in open(path, errors="replace"):
        body = ln.split("](biokanga) ", 1)[1] if "](biokanga) " in ln else ln
        if "Alignment of" in body and "completed" in body:
            on = True
        if on and ("Reporting of aligned result set" in body or "Exit code" in body):
            continue
        if on:
            keep.append(body.rstrip("\n"))
    return keep


@pytest.mark.parametrize("case,tag", RUNS)
def test_cli_outputs_match_reference(case, tag, golden_dir, tmp_path, src_case=None):
    assert os.path.exists(CLI), "bkx-align is not built"
    src = src_case or case               # genome + reads of another case (runs that only add options to its reads)
    run = gu.runs(case)[tag]
    sfx = gu.sfx_path(src, golden_dir)
    files = [os.path.join(gu.GOLD, src, f) for f in run["reads"]]  # .gz inputs are read directly
    base = [CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + run["args"]
    # -M0 CSV + log
    subprocess.run(base + ["-M0", "-o", str(tmp_path / "o.csv"), "-F", str(tmp_path / "o.log")], check=True,
                   stdout=subprocess.DEVNULL)
    ours = sorted(open(tmp_path / "o.csv").read().splitlines())
    ref = sorted(gzip.open(os.path.join(gu.GOLD, case, tag + ".csv.gz"), "rt").read().splitlines())
    assert ours == ref
    exp_log = open(os.path.join(gu.GOLD, case, tag + ".log")).read().splitlines()
    assert summary_block(tmp_path / "o.log") == exp_log
    # -M6 SAM: identical header, identical record set
    subprocess.run(base + ["-M6", "-o", str(tmp_path / "o.sam")], check=True, stdout=subprocess.DEVNULL)
    ours = open(tmp_path / "o.sam").read().splitlines()
    ref = gzip.open(os.path.join(gu.GOLD, case, tag + ".sam.gz"), "rt").read().splitlines()
    assert [x for x in ours if x.startswith("@")] == [x for x in ref if x.startswith("@")]
    assert sorted(x for x in ours if not x.startswith("@")) == sorted(x for x in ref if not x.startswith("@"))
    # accepted records come out in the reference's coordinate order
    def coords(lines, names):
        out = []
        for x in lines:
            c = x.split("\t")
            if not x.startswith("@") and not int(c[1]) & 4:
                out.append((names.index(c[2]), int(c[3])))
        return out
    names = [x.split("\t")[2][3:] for x in ref if x.startswith("@SQ")]
    co = coords(ours, names)
    assert co == sorted(co)


def test_cli_rejects_unsupported_and_bad_options(tmp_path, golden_dir):
    sfx = gu.sfx_path("tiny", golden_dir)
    rd = os.path.join(gu.GOLD, "tiny", "r100.fa.gz")
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-r2", "-R5"], capture_output=True, text=True)
    assert r.returncode != 0 and "not supported" in r.stderr
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-s99"], capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([CLI, "align", "-I", "/nonexistent.sfx", "-i", rd, "-o", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-x8"], capture_output=True, text=True)
    assert r.returncode != 0 and "0..7" in r.stderr
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-Zchr[1"], capture_output=True, text=True)
    assert r.returncode != 0     # malformed expression: refused like the reference's regcomp failure
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-M6", "-O", str(tmp_path / "st")], capture_output=True, text=True)
    assert r.returncode != 0 and "not available in '-M6'" in r.stderr
    for k, body in enumerate(("chrX,1,2,A\n", "chr1,5,2,A\n", "chr1,1,2,Q\n", "chr1,1,99999999,A\n", "chr1,1,2\n")):
        cf = tmp_path / ("bad%d.csv" % k)     # unknown chromosome / start > end / bad base letter / end beyond the sequence / too few fields
        cf.write_text(body)
        r = subprocess.run([CLI, "align", "-I", sfx, "-i", rd, "-o", str(tmp_path / "x"), "-5", str(cf)], capture_output=True, text=True)
        assert r.returncode != 0


def _lines(path):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        return f.read().splitlines()


@pytest.mark.parametrize("tag", ["m1", "m2", "m3", "m4", "m4t", "g0", "g1", "g2", "gz", "trim"])
def test_cli_output_formats_match_reference(tag, golden_dir, tmp_path):
    """CSV variants -M1..3, BED -M4 (default and -t title), FASTQ qualities -g0/-g1/-g2 in SAM, gzip output, end trims."""
    import json
    fdir = os.path.join(gu.GOLD, "formats")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path("tiny", golden_dir)
    out = tmp_path / run["out"]
    subprocess.run([CLI, "align", "-I", sfx, "-i", os.path.join(fdir, "q.fq.gz"), "-o", str(out)] + run["args"], check=True,
                   stdout=subprocess.DEVNULL)
    ours = _lines(out)
    gold = os.path.join(fdir, run["out"] if run["out"].endswith(".gz") else run["out"] + ".gz")
    ref = _lines(gold)
    if run["out"].endswith(".gz"):
        assert open(out, "rb").read(2) == b"\x1f\x8b"  # really gzip
    head = [x for x in ref if x.startswith("@") or x.startswith("track")]
    assert [x for x in ours if x.startswith("@") or x.startswith("track")] == head
    assert sorted(ours) == sorted(ref)


POST_TAGS = ["Z2", "Z2sam", "z13", "zZ", "ZZ", "x5", "x5m3", "x5bed", "x5sam", "x7", "x3Q1", "x6Z", "jJ", "jJr1", "x7r50", "x7r50sam"]


@pytest.mark.parametrize("tag", POST_TAGS)
def test_cli_post_alignment_passes_match_reference(tag, golden_dir, tmp_path):
    """Post-alignment host passes: chromosome filters -Z / -z, flank auto-trimming -x in every output
    format (trimmed coordinates, soft clips, eNARTrim), the -j / -J read reports; records, header and summary block."""
    import json
    fdir = os.path.join(gu.GOLD, "post")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path("tiny", golden_dir)
    reads = os.path.join(fdir, run.get("reads", "p.fa.gz"))
    subprocess.run([CLI, "align", "-I", sfx, "-i", reads, "-o", run["out"], "-F", "o.log"] + run["args"], check=True,
                   stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))
    head = [x for x in ref if x.startswith("@") or x.startswith("track")]
    assert [x for x in ours if x.startswith("@") or x.startswith("track")] == head
    assert sorted(ours) == sorted(ref)
    exp_log = [x for x in open(os.path.join(fdir, tag + ".log")).read().splitlines()
               if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp_log
    for a in run["args"]:          # -j / -J: same FASTA records (the reference's order among equal sort keys is unspecified)
        if a[:2] in ("-j", "-J"):
            name = a[2:]
            got = _lines(tmp_path / name)
            exp = _lines(os.path.join(fdir, name if name.endswith(".gz") else name + ".gz"))
            recs = lambda ls: sorted("\n".join(ls).split(">")[1:])
            assert recs(got) == recs(exp) and len(got) == len(exp)
            if name.endswith(".gz"):
                assert open(tmp_path / name, "rb").read(2) == b"\x1f\x8b"


@pytest.mark.parametrize("tag", ["k0", "k50", "k250", "k20sam", "k0x4"])
def test_cli_pcr_artefact_reduction_matches_reference(tag, golden_dir, tmp_path):
    """-k (ReducePCRduplicates): same summary block (incl. the removed count and the DP class) and the same alignments.  Which
    of several reads with identical sort keys is kept is left open by the reference's unstable sort, so rows are compared
    without read id / name (CSV) and on flag, chromosome, position, CIGAR and class (SAM)."""
    import json
    fdir = os.path.join(gu.GOLD, "dups")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path("tiny", golden_dir)
    subprocess.run([CLI, "align", "-I", sfx, "-i", os.path.join(fdir, "dup.fa.gz"), "-o", run["out"], "-F", "o.log"] + run["args"],
                   check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))

    def key(ln):
        if run["out"].endswith(".sam"):
            c = ln.split("\t")
            return ln if ln.startswith("@") else "\t".join([c[1], c[2], c[3], c[5]] + c[11:])
        return ",".join(ln.split(",")[1:13])
    assert sorted(map(key, ours)) == sorted(map(key, ref))
    exp_log = [x for x in open(os.path.join(fdir, tag + ".log")).read().splitlines()
               if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp_log


@pytest.mark.parametrize("tag", ["c5", "c5sam", "c5k", "c5pe", "c5pesam"])
def test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="constraints"):
    """-5 (LoadLociConstraints + IdentifyConstraintViolations): single-end, with -k / -x behind it, and paired-end runs."""
    import json
    fdir = os.path.join(gu.GOLD, case)
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path(run.get("index", "tiny"), golden_dir)
    files = [os.path.join(gu.GOLD, run.get("reads_dir", run.get("index", "tiny")), f) for f in run["reads"]]
    args = [a if not a.startswith("-5") else "-5" + os.path.join(fdir, a[2:]) for a in run["args"]]
    args = [os.path.join(fdir, a) if i and args[i - 1] == "-B" else a for i, a in enumerate(args)]   # -B <file of the case>
    subprocess.run([CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + args +
                   ["-o", run["out"], "-F", "o.log"], check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))
    if tag == "p6sam":  # -k behind it, SAM: compare on flag, chromosome, position, CIGAR and class
        def strip(ls):
            return sorted(x if x.startswith("@") else "\t".join([x.split("\t")[i] for i in (1, 2, 3, 5)] + x.split("\t")[11:]) for x in ls)
        assert strip(ours) == strip(ref)
    elif tag in ("c5k", "i4", "bv_x2k"):   # -k behind it: survivors among identical keys are unspecified -> compare without read id / name
        strip = lambda ls: sorted(",".join(x.split(",")[1:13]) for x in ls)
        assert strip(ours) == strip(ref)
    else:
        assert [x for x in ours if x.startswith(("@", "track"))] == [x for x in ref if x.startswith(("@", "track"))]
        assert sorted(ours) == sorted(ref)
    exp_log = [x for x in open(os.path.join(fdir, tag + ".log")).read().splitlines()
               if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp_log


@pytest.mark.parametrize("tag", ["s7", "s3pe", "pex0", "pex6", "p6", "p6sam", "p6pe", "p6pesam"])
def test_cli_read_sampling_matches_reference(tag, golden_dir, tmp_path):
    """-# (every Nth raw read / read pair of each file, taken before the length filter); pex*: -x in paired-end runs
    (trimmed POS / CIGAR / PNEXT / TLEN, nothing sloughed); p6*: -6 correction of 5' primer artefacts (p6pe*: behind -U2 / -U4 pairing)."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="sample")


@pytest.mark.parametrize("tag", ["o1", "o2", "o3pe"])
def test_cli_statistics_file_matches_reference(tag, golden_dir, tmp_path):
    """-O: the statistics CSV (insert-length histogram in paired-end runs, multi-hit distribution, bases and aligner induced
    substitutions per read offset and Phred band, substitutions per alignment, hits per target) is the reference's file."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="stats")
    assert _lines(tmp_path / "st.csv") == _lines(os.path.join(gu.GOLD, "stats", tag + ".st.csv.gz"))


@pytest.mark.parametrize("tag", ["i1", "i2", "i3", "i4", "i5"])
def test_cli_option_interplay_matches_reference(tag, golden_dir, tmp_path):
    """Combinations found by the randomised front-end comparison (tests/fuzz_host_cli.py): -r5 with -Z / -z (loci filtered as
    they are recorded, with the reference's compaction quirk), -r5 with BED (track line twice), -j / -J dropped under -r5,
    -O with BED (no substitution statistics)."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="interplay")
    st = os.path.join(gu.GOLD, "interplay", tag + ".st.csv.gz")
    if os.path.exists(st):
        assert _lines(tmp_path / "st.csv") == _lines(st)
    assert not os.path.exists(tmp_path / "n.fa") and not os.path.exists(tmp_path / "m.fa")


SIM_TAGS = ["se_g1", "se_g1bed", "se_g1sam", "se_g1x", "se_g3", "se_g3r1", "pe_U1", "pe_U2", "pe_U3sam", "lc_r5", "lc_r3", "piped",
            "pipedbed", "mix_a", "mix_b"]


@pytest.mark.parametrize("tag", SIM_TAGS)
def test_cli_simulated_read_truth_check_matches_reference(tag, golden_dir, tmp_path):
    """Reads written by the reference's own `simreads` carry their origin in the descriptor; `align` checks every accepted
    alignment against it (ReportAlignStats, Aligner.cpp:3556-3657): the summary gains "There are N (a 2 edge, b 1 edge) high
    confidence aligned simulated reads with M misaligned" and misaligned reads are written as "iar" (CSV, BED).  SE / PE,
    behind -x / -r1 / -r3 / -r5 / -Z, chromosome names holding '|' (second descriptor form), and read sets that mix simulated
    and plain reads (the first accepted read decides; the first one that does not parse ends the checking)."""
    import json
    import shutil
    fdir = os.path.join(gu.GOLD, "simreads")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    if run["index"] in gu.CASES:
        sfx = gu.sfx_path(run["index"], golden_dir)
    else:
        sfx = os.path.join(str(golden_dir), run["index"] + ".sfx")
        if not os.path.exists(sfx):
            with gzip.open(os.path.join(fdir, run["index"] + ".sfx.gz"), "rb") as a, open(sfx, "wb") as b:
                shutil.copyfileobj(a, b)
    files = [os.path.join(fdir, f) for f in run["reads"]]
    subprocess.run([CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + run["args"] +
                   ["-o", run["out"], "-F", "o.log"], check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))
    assert [x for x in ours if x.startswith(("@", "track"))] == [x for x in ref if x.startswith(("@", "track"))]
    if tag == "lc_r5":   # one record per locus, numbered in read order: line for line
        assert {ln.split(",", 1)[0]: ln for ln in ours} == {ln.split(",", 1)[0]: ln for ln in ref}
    assert sorted(ours) == sorted(ref)
    if tag in ("lc_r5", "lc_r3", "se_g3", "piped", "pipedbed"):   # the fixtures that hold misaligned reads
        assert any('"iar"' in x or "\tiar\t" in x for x in ref)
    exp_log = [x for x in open(os.path.join(fdir, tag + ".log")).read().splitlines()
               if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp_log


def test_cli_parameter_file_long_options_and_wildcards_match_reference(golden_dir, tmp_path):
    """Option grammar (kanga.cpp:194-298, Utility.cpp:793-912): "@file" parameter files (several options per line, comment
    lines, a quoted value), the long option names (--name=value and --name value), bundled flags, and a single-end input
    specification with wildcards, whose matches load in case-insensitive name order (the read ids of the CSV pin it)."""
    import shutil
    fdir = os.path.join(gu.GOLD, "grammar")
    sfx = gu.sfx_path("tiny", golden_dir)
    for nm in ("A_part.fa", "b_part.fa", "C_part.fa"):
        with gzip.open(os.path.join(fdir, nm + ".gz"), "rb") as a, open(tmp_path / nm, "wb") as b:
            shutil.copyfileobj(a, b)
    shutil.copyfile(os.path.join(fdir, "params.txt"), tmp_path / "params.txt")
    subprocess.run([CLI, "align", "@params.txt", "-I", sfx, "--out", "g1.csv", "--log=g1.log"], check=True, stdout=subprocess.DEVNULL,
                   cwd=tmp_path)
    ref = _lines(os.path.join(fdir, "g1.csv.gz"))
    assert sorted(_lines(tmp_path / "g1.csv")) == sorted(ref)
    assert summary_block(tmp_path / "g1.log") == open(os.path.join(fdir, "g1.log")).read().splitlines()
    # the same run spelled with short options, the three files named one by one in load order
    subprocess.run([CLI, "align", "-I", sfx, "-i", "A_part.fa", "-i", "b_part.fa", "-iC_part.fa", "-s", "3", "-M0", "-o", "g2.csv"],
                   check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    assert _lines(tmp_path / "g2.csv") == _lines(tmp_path / "g1.csv")
    # bundled literal flags (-XE is -X -E), -q / -w / -W accepted (the summary database itself is not written)
    subprocess.run([CLI, "align", "-I", sfx, "-i", "?_part.fa", "-Xs3", "-M0", "-o", "g3.csv", "-q", "sum.db", "-w", "exp", "-W", "descr"],
                   check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    assert _lines(tmp_path / "g3.csv") == _lines(tmp_path / "g1.csv")
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", "A_part.fa", "-o", "x.csv", "-q", "sum.db"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "-w<str>" in r.stderr
    r = subprocess.run([CLI, "align", "-I", sfx, "-i", "nomatch*.fa", "-o", "x.csv"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "Unable to glob" in r.stdout
    r = subprocess.run([CLI, "align", "@missing.txt", "-I", sfx, "-o", "x.csv"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "Unable to open options file" in r.stdout


BEST_TAGS = ["n5_r5", "n5_r5sam", "n3_r5s8", "n8_r5Q1", "n2_r5s0", "n4_r1", "n5_r3", "n6_r4"]


@pytest.mark.parametrize("tag", BEST_TAGS)
def test_cli_best_matches_match_reference(tag, golden_dir, tmp_path):
    """-N (CSfxArrayV3::LocateBestMatches, SfxArrayV2.cpp:6654-7019): the -R<n> loci with the fewest mismatches from one
    un-staged pass, behind -r1 (distribution only), -r3 / -r4 (clustering) and -r5 (every locus a record, numbered in read
    order: compared line for line); -s0, -e2, -Q1 among the runs."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="bestmatches")
    import json
    run = json.load(open(os.path.join(gu.GOLD, "bestmatches", "runs.json")))[tag]
    if "-r5" in run["args"] and run["out"].endswith(".csv"):
        ours = {ln.split(",", 1)[0]: ln for ln in _lines(tmp_path / run["out"])}
        ref = {ln.split(",", 1)[0]: ln for ln in _lines(os.path.join(gu.GOLD, "bestmatches", run["out"] + ".gz"))}
        assert ours == ref


PRIORITY_TAGS = ["b_s3", "bv_s3", "bv_s5e2", "b_csv_sam", "bv_deep", "b_z", "bv_x2k", "bv_r1", "b_r1x", "bv_r3", "b_r4x", "bv_r5", "b_r5x", "bv_r5n", "b_pe", "bv_pe", "b_pe3", "bv_pe4"]


@pytest.mark.parametrize("tag", PRIORITY_TAGS)
def test_cli_priority_regions_match_reference(tag, golden_dir, tmp_path):
    """-B / -V (Aligner.cpp:9102-9186, 4126-4186): a read whose first search -- with room for 10 more loci -- ends with exactly
    one locus inside a region of the BED file is accepted there as unique; without -V the accepted alignments outside every
    region become PR.  BED with tabs / blanks / commas, comments, a header line, upper-case and unknown chromosome names,
    overlapping features; behind it -Z, -x, -k; -e2; -M6; in front of every multi-loci mode (up to -R loci inside the regions
    stand for the read), with -X and -N; in front of the pairing of paired-end runs (-U1..4, orphan recovery, -Z, -x)."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="priority")


CONTAM_TAGS = ["h_s3", "h_trim", "h_sam", "h_fq", "h_pe", "h_pesam", "h_x", "hv_s3", "hv_sam", "hv_pe"]


@pytest.mark.parametrize("tag", CONTAM_TAGS)
def test_cli_adaptor_trimming_matches_reference(tag, golden_dir, tmp_path):
    """-H (CContaminants, Contaminants.cpp:204-431, 1227-1310; Aligner.cpp:11036-11084): adaptor tails at the read ends are
    trimmed off at load -- every overlay code (@1, @3, @1234, reverse complements, none, PE2 only), an N in an adaptor, the
    4-base minimum; reads carrying tails of 1..30 bases with and without a substitution; with fixed trims and a length filter,
    FASTQ qualities, SAM, paired ends with orphan recovery, sampling, -x and -Z behind it; hv_*: vector sequences ('&'
    codes) that reads cut out of them -- sense or antisense, 0..6 substitutions -- lie inside."""
    import json
    fdir = os.path.join(gu.GOLD, "contam")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path("tiny", golden_dir)
    files = [os.path.join(fdir, f) for f in run["reads"]]
    args = [os.path.join(fdir, a) if i and run["args"][i - 1] == "-H" else a for i, a in enumerate(run["args"])]
    subprocess.run([CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + args +
                   ["-o", run["out"], "-F", "o.log"], check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))
    assert [x for x in ours if x.startswith("@")] == [x for x in ref if x.startswith("@")]
    assert sorted(ours) == sorted(ref)
    exp_log = [x for x in open(os.path.join(fdir, tag + ".log")).read().splitlines()
               if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp_log


@pytest.mark.parametrize("tag", ["e_m0", "e_m6", "e_m4x", "e_pe", "e_r5", "e_r4", "e_r1"])
def test_cli_run_without_a_read_matches_reference(tag, golden_dir, tmp_path):
    """No read survives the load filters: like the reference the front end goes on and reports an empty run (average length 0
    with a minimum of -1, one unprocessed record in the class summary, an empty record in -M6)."""
    test_cli_loci_base_constraints_match_reference(tag, golden_dir, tmp_path, case="empty")


MANY_TAGS = ["r5_R300X", "r5_R500", "r5_R100N", "r4_R200X", "r1_R500"]


@pytest.mark.parametrize("tag", MANY_TAGS)
def test_cli_hundreds_of_loci_per_read_match_reference(tag, golden_dir, tmp_path):
    """-R up to the reference's 500 (cMaxMultiHits) behind -r1 / -r4 / -r5, with -X and -N: 60 bp reads inside the 64-base unit
    of the `repeats` genome carry ~2850 equally good loci, the ones inside the 90-base unit dozens to hundreds."""
    import json
    fdir = os.path.join(gu.GOLD, "manyloci")
    run = json.load(open(os.path.join(fdir, "runs.json")))[tag]
    sfx = gu.sfx_path("repeats", golden_dir)
    subprocess.run([CLI, "align", "-I", sfx, "-i", os.path.join(fdir, run["reads"][0])] + run["args"] + ["-o", run["out"], "-F", "o.log"],
                   check=True, stdout=subprocess.DEVNULL, cwd=tmp_path)
    ours, ref = _lines(tmp_path / run["out"]), _lines(os.path.join(fdir, run["out"] + ".gz"))
    if "-r5" in run["args"]:   # numbered in read order: line for line
        assert {ln.split(",", 1)[0]: ln for ln in ours} == {ln.split(",", 1)[0]: ln for ln in ref}
    assert sorted(ours) == sorted(ref)
    assert summary_block(tmp_path / "o.log") == open(os.path.join(fdir, tag + ".log")).read().splitlines()


def _bgzf_blocks(raw):
    import struct
    o, out = 0, []
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04"
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        out.append((bsize, struct.unpack_from("<I", raw, o + bsize - 4)[0]))
        o += bsize
    return out


BAM_RUNS = [("bam5", ["-s3", "-M5"], "out5.bam"), ("bam6", ["-s3", "-M6", "-g0"], "out6.bam"), ("bamQ2", ["-s3", "-M6", "-Q2"], "out62.bam"),
            ("x5bam", ["-s5", "-M5", "-x5"], "x5.bam"), ("x5bam6", ["-s5", "-M6", "-x5", "-Zchr2"], "x56.bam")]


@pytest.mark.parametrize("tag,args,out", BAM_RUNS)
def test_cli_bam_and_bai_match_reference(tag, args, out, golden_dir, tmp_path):
    """BAM (BGZF) + BAI written for an output name ending in .bam: byte-identical to the reference's files
    (the read sets have no two reads at the same locus, so sort ties cannot reorder records); the x5* runs carry
    soft clips from -x trimming and records filtered by -Z."""
    fdir = os.path.join(gu.GOLD, "post" if tag.startswith("x5") else "formats")
    sfx = gu.sfx_path("tiny", golden_dir)
    reads = os.path.join(fdir, "p.fa.gz" if tag.startswith("x5") else "qu.fq.gz")
    subprocess.run([CLI, "align", "-I", sfx, "-i", reads, "-o", str(tmp_path / out)] + args, check=True,
                   stdout=subprocess.DEVNULL)
    ours, ref = open(tmp_path / out, "rb").read(), open(os.path.join(fdir, out), "rb").read()
    assert gzip.decompress(ours) == gzip.decompress(ref)
    assert _bgzf_blocks(ours) == _bgzf_blocks(ref)
    assert ours == ref
    assert open(str(tmp_path / out) + ".bai", "rb").read() == open(os.path.join(fdir, out + ".bai"), "rb").read()


def test_sort_hits_orders_like_the_host_comparator():
    """bkx_sort_hits (device radix passes) == a host sort on the SortHitMatch key tuple, ties by record index."""
    import numpy as np
    from biokanga_b200 import abi
    rng = np.random.default_rng(5)
    n = 300_000
    r = np.zeros(n, dtype=abi.RESULT_DTYPE)
    r["nar"] = rng.choice([1, 1, 1, 2, 3, 4, 5, 13, 15], n)
    r["num_hits"] = np.where(r["nar"] == 1, 1, rng.integers(0, 3, n))
    r["chrom_id"] = rng.integers(1, 40, n)
    r["match_loci"] = rng.integers(0, 2000, n)          # many ties on purpose
    r["match_len"] = rng.choice([100, 150], n)
    r["strand"] = rng.choice([ord("+"), ord("-")], n)
    r["low_mm"] = rng.integers(0, 4, n)
    order = bkx.sort_hits(r)
    uniq = r["num_hits"] == 1
    keys = np.stack([r["nar"].astype(np.int64), (~uniq).astype(np.int64), np.where(uniq, 0, r["num_hits"]).astype(np.int64),
                     np.where(uniq, r["chrom_id"], 0).astype(np.int64), np.where(uniq, r["match_loci"], 0).astype(np.int64),
                     np.where(uniq, r["match_len"], 0).astype(np.int64), np.where(uniq, r["strand"], 0).astype(np.int64),
                     np.where(uniq, r["low_mm"], 0).astype(np.int64), np.arange(n)], axis=0)
    exp = np.lexsort(keys[::-1])
    assert np.array_equal(order, exp.astype(np.uint32))


@pytest.mark.parametrize("case,tag", [("tiny", "r100_s3"), ("tiny", "mixed_s3"), ("tiny", "pe_U1"), ("repeats", "r60_s5")])
def test_cli_chunked_parallel_parse_gives_the_same_files(case, tag, golden_dir, tmp_path):
    """The read files are split at record boundaries and parsed by several threads; forced here onto tiny files
    (chunks of ~3 kB, 7 threads): CSV rows identical to the reference's."""
    run = gu.runs(case)[tag]
    sfx = gu.sfx_path(case, golden_dir)
    files = [os.path.join(gu.GOLD, case, f) for f in run["reads"]]
    base = [CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + run["args"]
    env = dict(os.environ, BKX_PARSE_MIN_CHUNK="3000")
    subprocess.run(base + ["-T7", "-M0", "-o", str(tmp_path / "o.csv"), "-F", str(tmp_path / "o.log")], check=True,
                   stdout=subprocess.DEVNULL, env=env)
    ours = sorted(open(tmp_path / "o.csv").read().splitlines())
    ref = sorted(gzip.open(os.path.join(gu.GOLD, case, tag + ".csv.gz"), "rt").read().splitlines())
    assert ours == ref
    exp_log = open(os.path.join(gu.GOLD, case, tag + ".log")).read().splitlines()
    assert summary_block(tmp_path / "o.log") == exp_log


@pytest.mark.parametrize("tag", ["r5_R5_s3", "r5_R3_X_s3", "r5_R8_s5_e2"])
def test_cli_all_loci_mode_sam_log_and_numbering(tag, golden_dir, tmp_path):
    """-r5 with -M6: reads without a locus stay in the record set, so the summary differs from the -M0 run -- compared
    with the reference's own -M6 log; and the CSV keeps the reference's record numbering (one record per locus, in
    read order) when compared line by line, not just as a set."""
    run = gu.runs("lowcopy")[tag]
    sfx = gu.sfx_path("lowcopy", golden_dir)
    rd = os.path.join(gu.GOLD, "lowcopy", run["reads"][0])
    base = [CLI, "align", "-I", sfx, "-i", rd] + run["args"]
    subprocess.run(base + ["-M6", "-o", str(tmp_path / "o.sam"), "-F", str(tmp_path / "o.log")], check=True,
                   stdout=subprocess.DEVNULL)
    exp = [x for x in open(os.path.join(gu.GOLD, "lowcopy", tag + ".sam.log")).read().splitlines()
           if not x.startswith(("Sorting alignments", "Header written", "Reported SAM", "Completed reporting SAM"))]
    assert summary_block(tmp_path / "o.log") == exp
    subprocess.run(base + ["-M0", "-o", str(tmp_path / "o.csv")], check=True, stdout=subprocess.DEVNULL)
    ours = {ln.split(",", 1)[0]: ln for ln in open(tmp_path / "o.csv").read().splitlines()}
    ref = {ln.split(",", 1)[0]: ln for ln in gzip.open(os.path.join(gu.GOLD, "lowcopy", tag + ".csv.gz"), "rt").read().splitlines()}
    assert ours == ref
