"""The native TSV reader / writer and whitelist reader (badger_b200.tsvio -> bdg_tsv_* / bdg_lines16_*) against pandas,
i.e. against what the reference does at badger.py:82-111 and barcode_graph.py:388-410, and the CLI route built on them
against the golden output files of the unmodified reference.  Host code: runs without a GPU (the device operators of
the CLI test are the oracle-backed stand-ins of tests/cpu_ops.py)."""
import importlib.util
import logging
import os

import numpy as np
import pandas as pd
import pytest

import badger_b200
from badger_b200 import tsvio
from oracle import oracle as orc

import cpu_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = "#read_id\tbarcode\tUMI\tBC_score\tvalid_UMI\tstrand\tpolyT_start\tR1_end\n"


def reference_ingest(path, bc_len=16):
    """badger.py:91-111 verbatim in spirit: what the graph gets (`barcodes`, after barcode_graph.py:195-197's length rule)
    and what the writer gets (`read_assignment`)."""
    reads = pd.read_csv(path, sep="\t")
    ids = reads["#read_id"].tolist()
    observed = reads["barcode"].fillna('*').tolist()
    barcodes = reads["barcode"].dropna()
    barcodes = barcodes[barcodes != "*"]
    barcodes = barcodes[barcodes != "barcode"].tolist()
    graph = [b[:-1] if len(b) == bc_len + 1 else b for b in barcodes if len(b) in (bc_len, bc_len + 1)]
    read_assignment = []
    for i in range(len(ids)):
        if ids[i] != "#read_id":
            o = observed[i]
            if o != "barcode":
                if len(o) == bc_len + 1:
                    o = o[:-1]
                read_assignment.append((ids[i], o))
    return graph, read_assignment


def native_ingest(path):
    with tsvio.ExtractionTsv(path) as t:
        has, emit = t.has_barcode, t.emitted
        recs = [bytes(r).decode() for r in t.seqs16]
        graph = [recs[i] for i in np.nonzero(has)[0]]
        return graph, [(i, recs[i] if has[i] else None) for i in np.nonzero(emit)[0]], t.rows


def random_tsv(rng, n, tricky=True):
    lines = [HDR]
    for i in range(n):
        bc = "".join(rng.choice(list("ACGT"), 16))
        r = rng.random()
        if tricky:
            if r < 0.05: bc = "*"
            elif r < 0.08: bc = ""
            elif r < 0.10: bc = rng.choice(["NA", "nan", "NULL", "N/A", "None", "<NA>", "n/a", "#N/A"])
            elif r < 0.15: bc = bc + "ACGT"[rng.integers(4)]           # 17-mer: loses its last base
            elif r < 0.17: bc = bc[:rng.integers(1, 16)]               # too short: never in the graph, written as '*'
            elif r < 0.19: bc = bc + "ACGTAC"[:rng.integers(2, 6)]     # too long
            elif r < 0.20: bc = "barcode"
        rid = "read_%d" % i if rng.random() > 0.02 or not tricky else rng.choice(["#read_id", " lead", "trail ", "a b", "x#y", "0a1f-77", "r/1"])
        fields = [rid, bc, "ACGTACGTACGT", "16", "True", "+", "40", "21"]
        if tricky and rng.random() < 0.03:
            fields = fields[:rng.integers(1, 8)]                        # short row: pandas pads with NaN
        lines.append("\t".join(fields) + "\n")
        if tricky and rng.random() < 0.02:
            lines.append(rng.choice(["\n", "   \n", HDR]))              # blank lines are skipped, repeated headers dropped
    text = "".join(lines)
    if tricky and rng.random() < 0.5:
        text = text.rstrip("\n")                                        # no newline at the end of the file
    return text


def test_reader_matches_pandas_on_golden(gold_pipeline):
    path = gold_pipeline["dir"] + "/reads.tsv"
    graph, ra = reference_ingest(path)
    ngraph, nra, rows = native_ingest(path)
    assert ngraph == graph and len(nra) == len(ra) <= rows
    assert [o for _, o in nra] == [o if len(o) == 16 and o != "barcode" else None for _, o in ra]


@pytest.mark.parametrize("seed", range(6))
def test_reader_matches_pandas_on_tricky_files(tmp_path, seed):
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "reads.tsv")
    with open(path, "w") as fh:
        fh.write(random_tsv(rng, 3000 if seed else 150000))             # seed 0: large enough for several parser threads
    graph, ra = reference_ingest(path)
    ngraph, nra, _ = native_ingest(path)
    assert ngraph == graph
    assert len(nra) == len(ra)
    # the writer only needs: which rows are written, and the 16-mer to look up ('*', NaN, other lengths all end as '*')
    want = [o if len(o) == 16 and o != "barcode" else None for _, o in ra]
    assert [o for _, o in nra] == want
    # writer: ids byte for byte as to_csv writes them
    with tsvio.ExtractionTsv(path) as t:
        centre = np.full(t.rows, tsvio.NONE, np.uint64)
        pick = t.has_barcode & (rng.random(t.rows) < 0.7)
        vals = rng.integers(0, 1 << 32, t.rows, dtype=np.uint64)
        centre[pick] = vals[pick]
        out = str(tmp_path / "native.tsv")
        t.write(out, centre)
        emitted = np.nonzero(t.emitted)[0]
    res = ["*" if centre[i] == tsvio.NONE else orc.unrank(int(centre[i])) for i in emitted]
    pd.DataFrame({"readID": [r[0] for r in ra], "barcode": res}).to_csv(str(tmp_path / "pandas.tsv"), sep="\t", index=False)
    assert open(out, "rb").read() == open(str(tmp_path / "pandas.tsv"), "rb").read()


def test_reader_fuzz_accepts_only_what_it_reads_like_pandas(tmp_path):
    """Small hostile files (random fields, stray tabs, numbers, NA strings, spaces, '#'): the native reader either declines
    or agrees with pandas on the graph's barcodes, on the written rows and - through the writer - on the ids byte for byte."""
    import warnings
    rng = np.random.default_rng(2)
    alph = list("ACGTN*ab1.e-   #/_\t\n") + ["NA", "nan", "barcode", "#read_id", "read_", "ACGTACGTACGTACGT", "ACGTACGTACGTACGTA", "True", "12", ""]
    path, out_n, out_p = str(tmp_path / "f.tsv"), str(tmp_path / "n.tsv"), str(tmp_path / "p.tsv")
    accepted = 0
    for trial in range(1200):
        cols = ["#read_id", "barcode", "UMI"]
        if rng.random() < 0.2:
            rng.shuffle(cols)
        lines = ["\t".join(cols)]
        for r in range(int(rng.integers(1, 8))):
            if rng.random() < 0.6:
                rid = "r%d" % r if rng.random() < 0.8 else "".join(rng.choice(alph, int(rng.integers(0, 4))))
                bc = str(rng.choice(["ACGTACGTACGTACGT", "ACGTACGTACGTACGTA", "*", "", "NA", "ACGT", "barcode", "TTTTACGTACGTACGT "])) \
                    if rng.random() < 0.8 else "".join(rng.choice(alph, int(rng.integers(0, 4))))
                f = [rid, bc, "x"]
                lines.append("\t".join(f[:int(rng.integers(1, 4))] if rng.random() < 0.2 else f))
            else:
                lines.append("".join(rng.choice(alph, int(rng.integers(0, 6)))))
        with open(path, "w") as fh:
            fh.write("\n".join(lines) + ("\n" if rng.random() < 0.5 else ""))
        try:
            ngraph, nra, _ = native_ingest(path)
        except tsvio.Unsupported:
            continue
        accepted += 1
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            graph, ra = reference_ingest(path)                 # must not raise on a file the native reader accepted
        assert ngraph == graph, lines
        assert [o for _, o in nra] == [o if isinstance(o, str) and len(o) == 16 and o != "barcode" else None for _, o in ra], lines
        with tsvio.ExtractionTsv(path) as t:
            t.write(out_n, np.full(t.rows, tsvio.NONE, np.uint64))
        pd.DataFrame({"readID": [r[0] for r in ra], "barcode": ["*"] * len(ra)}).to_csv(out_p, sep="\t", index=False)
        assert open(out_n, "rb").read() == open(out_p, "rb").read(), lines
    assert accepted > 150


@pytest.mark.parametrize("body,why", [
    ('r1\t"ACGTACGTACGTACGT"\tx\n', "quote"),
    ("r1\tACGTACGTACGTACGT\tx\r\n", "carriage"),
    ("r1\tACGTACGTACGTACGT\tx\ty\n", "more fields"),
    ("123\tACGTACGTACGTACGT\tx\n", "read id"),
    ("1e5\tACGTACGTACGTACGT\tx\n", "read id"),
    ("NA\tACGTACGTACGTACGT\tx\n", "read id"),
    ("\tACGTACGTACGTACGT\tx\n", "read id"),
    ("True\tACGTACGTACGTACGT\tx\n", "read id"),
    ("r1\t12345\tx\n", "barcode field"),
    ("r\xc3\xa9\tACGTACGTACGTACGT\tx\n", "non-ASCII"),
    ("r1\t*\tx\n", "no 16/17"),
])
def test_reader_declines_what_pandas_treats_specially(tmp_path, body, why):
    path = str(tmp_path / "odd.tsv")
    with open(path, "wb") as fh:
        fh.write(("#read_id\tbarcode\tUMI\n" + "r0\tACGTACGTACGTACGA\tx\n" * (0 if why == "no 16/17" else 1) + body).encode("latin-1"))
    with pytest.raises(tsvio.Unsupported) as e:
        tsvio.ExtractionTsv(path)
    assert why in str(e.value)


def test_reader_header_and_io_errors(tmp_path):
    p = str(tmp_path / "nohdr.tsv")
    open(p, "w").write("id\tbc\nr1\tACGTACGTACGTACGT\n")
    with pytest.raises(tsvio.Unsupported):
        tsvio.ExtractionTsv(p)
    open(p, "w").write("")
    with pytest.raises(tsvio.Unsupported):
        tsvio.ExtractionTsv(p)
    with pytest.raises(badger_b200.BadgerB200Error) as e:
        tsvio.ExtractionTsv(str(tmp_path / "missing.tsv"))
    assert e.value.code == -7
    # columns in another order, barcode column last
    open(p, "w").write("UMI\tbarcode\t#read_id\nx\tACGTACGTACGTACGT\tr1\ny\t*\tr2\nz\tACGTACGTACGTACGTA\n")
    # the third row has no #read_id field: pandas reads NaN there, the native reader declines the file
    with pytest.raises(tsvio.Unsupported):
        tsvio.ExtractionTsv(p)
    open(p, "w").write("UMI\tbarcode\t#read_id\nx\tACGTACGTACGTACGT\tr1\ny\t*\tr2\n")
    graph, ra = reference_ingest(p)
    ngraph, nra, rows = native_ingest(p)
    assert ngraph == graph == ["ACGTACGTACGTACGT"] and rows == 2 and [o for _, o in nra] == ["ACGTACGTACGTACGT", None]


def test_whitelist_records(tmp_path, gold_pipeline):
    path = gold_pipeline["dir"] + "/whitelist.txt"
    with open(path) as fh:
        want = set(fh.read().split("\n"))
    got = {bytes(r).decode() for r in tsvio.whitelist_records(path)}
    assert got == {w for w in want if len(w) == 16}
    p = str(tmp_path / "wl.txt")
    open(p, "wb").write(b"ACGTACGTACGTACGT\r\nACGTACGTACGTACG\nTTTTACGTACGTACGTA\n\nGGGGACGTACGTACGT-1\nCCCCACGTACGTACGT")
    with open(p) as fh:
        want = {w for w in fh.read().split("\n") if len(w) == 16}
    assert {bytes(r).decode() for r in tsvio.whitelist_records(p)} == want == {"ACGTACGTACGTACGT", "CCCCACGTACGTACGT"}
    open(p, "wb").write(b"")
    assert tsvio.whitelist_records(p).shape == (0, 16)


def _cli():
    spec = importlib.util.spec_from_file_location("badger_cli", os.path.join(ROOT, "badger.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    return cli


def _run(cli, argv, capsys):
    cli.main(argv)
    for h in list(logging.getLogger("BarcodeGraph").handlers):
        logging.getLogger("BarcodeGraph").removeHandler(h)
    out = capsys.readouterr().out
    return [ln for ln in out.splitlines() if ln.strip().lstrip("-").isdigit()], out


def test_cli_native_route_golden(gold_pipeline, monkeypatch, tmp_path, capsys):
    """badger.py through the native reader / array pipeline / native writer: the reference's output file byte for byte,
    the same digits on stdout, and the same results as the pandas + dict route for --high_sens and --true_barcodes."""
    cpu_ops.install(monkeypatch)
    cli = _cli()
    monkeypatch.setattr(cli, "init", lambda *a, **k: 1)
    g = gold_pipeline
    base = ["-r", g["dir"] + "/reads.tsv", "-d", "tenX_v3", "-t", str(g["t"]), "--n_cells", str(g["n_cells"]), "-i", str(g["interval"])]
    wl = ["-l", g["dir"] + "/whitelist.txt"]
    out = str(tmp_path / "NATIVE")
    tail, text = _run(cli, base + wl + ["-o", out], capsys)
    assert "reading through pandas" not in text and "k: 6" in text
    assert open(out + "_output_file.tsv", "rb").read() == open(g["dir"] + "/expected_output_file.tsv", "rb").read()
    assert tail == g["stdout_tail"]
    tb = str(tmp_path / "true.tsv")
    with open(tb, "w") as fh:
        for c in g["centres"][:60]:
            fh.write(orc.unrank(c) + "-1\n")
    for extra in (["-hs"] + wl, ["--true_barcodes", tb], ["--true_barcodes", tb, "-hs"], []):
        a, b = str(tmp_path / "A"), str(tmp_path / "B")
        tail_a, _ = _run(cli, base + extra + ["-o", a], capsys)
        tail_b, _ = _run(cli, base + extra + ["-o", b, "--no_native_io"], capsys)
        assert open(a + "_output_file.tsv", "rb").read() == open(b + "_output_file.tsv", "rb").read(), extra
        assert tail_a == tail_b, extra


@pytest.mark.parametrize("seed", range(4))
def test_cli_routes_agree_on_random_inputs(seed, monkeypatch, tmp_path, capsys):
    """Array route (native reader, assign_packed with --high_sens resolved on node positions, native writer) against the
    pandas + dict route on fresh random inputs: thresholds 1 and 2, with / without whitelist, --high_sens, --true_barcodes,
    17-mers in the file, dense clusters that produce same-round conflicts (evictions)."""
    from badger_b200 import synth
    cpu_ops.install(monkeypatch)
    cli = _cli()
    monkeypatch.setattr(cli, "init", lambda *a, **k: 1)
    rng = synth.rng_for(500 + seed)
    wl = synth.make_whitelist(800, rng)
    cells = synth.pick_cells(wl, 40, rng)
    if seed % 2:                                            # near-identical cells: nodes claimed by two centres in one round
        cells[1::2] = cells[::2] ^ np.uint32(1 << int(rng.integers(0, 32)))
    obs, valid = synth.simulate_reads(cells, 1500, 0.07, rng)
    tsv, wlf, tb = str(tmp_path / "r.tsv"), str(tmp_path / "wl.txt"), str(tmp_path / "tb.tsv")
    synth.write_extraction_tsv(tsv, obs, valid, rng, extra17_frac=0.2)
    synth.write_whitelist(wlf, wl)
    with open(tb, "w") as fh:
        for c in cells[:25].tolist():
            fh.write(orc.unrank(int(c)) + "-1\n")
    t = 1 + seed % 2
    base = ["-r", tsv, "-d", "10x", "-t", str(t), "--n_cells", "40"]
    for extra in (["-l", wlf], ["-l", wlf, "-hs"], ["--true_barcodes", tb, "-hs"], ["-hs"]):
        a, b = str(tmp_path / "A"), str(tmp_path / "B")
        tail_a, text = _run(cli, base + extra + ["-o", a], capsys)
        assert "reading through pandas" not in text
        tail_b, _ = _run(cli, base + extra + ["-o", b, "--no_native_io"], capsys)
        assert open(a + "_output_file.tsv", "rb").read() == open(b + "_output_file.tsv", "rb").read(), extra
        assert tail_a == tail_b, extra


def test_cli_falls_back_to_pandas_when_the_reader_declines(gold_pipeline, monkeypatch, tmp_path, capsys):
    cpu_ops.install(monkeypatch)
    cli = _cli()
    monkeypatch.setattr(cli, "init", lambda *a, **k: 1)
    g = gold_pipeline
    odd = str(tmp_path / "reads_crlf.tsv")
    with open(g["dir"] + "/reads.tsv", "rb") as fh:
        open(odd, "wb").write(fh.read().replace(b"\n", b"\r\n"))
    out = str(tmp_path / "OUT")
    tail, text = _run(cli, ["-r", odd, "-l", g["dir"] + "/whitelist.txt", "-d", "10x", "-t", str(g["t"]), "--n_cells", str(g["n_cells"]),
                            "-i", str(g["interval"]), "-o", out], capsys)
    assert "reading through pandas" in text
    assert open(out + "_output_file.tsv", "rb").read() == open(g["dir"] + "/expected_output_file.tsv", "rb").read()
    assert tail == g["stdout_tail"]


def test_cli_native_route_keyerror_on_non_acgt(monkeypatch, tmp_path):
    cpu_ops.install(monkeypatch)
    cli = _cli()
    monkeypatch.setattr(cli, "init", lambda *a, **k: 1)
    p = str(tmp_path / "n.tsv")
    open(p, "w").write("#read_id\tbarcode\nr1\tACGTACGTACGTACGT\nr2\tACGTACGTNCGTACGT\n")
    with pytest.raises(KeyError) as e:
        cli.main(["-r", p, "-d", "10x", "-o", str(tmp_path / "O")])
    for h in list(logging.getLogger("BarcodeGraph").handlers):
        logging.getLogger("BarcodeGraph").removeHandler(h)
    assert e.value.args[0] == "N"
