"""Panel YAML loading (sharkmer_b200/panels.py) against the reference's unit tests
(src/pcr/preconfigured.rs:540-728) and its test fixture's shape (tests/fixtures/test_panel.yaml)."""
import pytest

from sharkmer_b200 import panels
from sharkmer_b200.primers import PCRParams

FIXTURE = """name: test_panel
description: "Test panel for URL loading (copy of teleostei 18S)"
primers:
  - gene: "18S"
    forward_seq: "TAACATATGCTTGTCTCAAAG"
    reverse_seq: "CCTGTATTGTTATTTTTCGTCAC"
    min_length: 300
    max_length: 700
    min_count: 2
    mismatches: 2
    trim: 15
    citation: "Test citation"
    notes: "test"
"""


def test_kats(tmp_path):
    assert panels.is_url("https://example.com/panel.yaml") and panels.is_url("http://example.com/panel.yaml")
    assert not panels.is_url("/path/to/panel.yaml") and not panels.is_url("relative/panel.yaml") and not panels.is_url("panel.yaml")
    assert panels.derive_gene_name("CO1") == "CO1" and panels.derive_gene_name("18S", "V9") == "18S-V9"
    assert panels.derive_gene_name("CO1", None, 2) == "CO1_2" and panels.derive_gene_name("18S", "V5-V7", 1) == "18S-V5-V7_1"
    for gene in ("Cyt-b", "CO-1"):
        with pytest.raises(panels.PanelError):
            panels.validate_gene_chars(gene, True)
    for gene in ("psbA-trnH", "trnL-F", "Cyt-b", "CO1", "18S", "5.8S"):
        panels.validate_gene_chars(gene, False)
    for has_region in (False, True):
        with pytest.raises(panels.PanelError):
            panels.validate_gene_chars("18S_rRNA", has_region)
    with pytest.raises(panels.PanelError):
        panels.validate_region_chars("V5_V7")
    panels.validate_region_chars("V5-V7")
    panels.validate_region_chars("V9")
    f = tmp_path / "test_panel.yaml"
    f.write_text(FIXTURE)
    prm = panels.load_panel_file(str(f))
    assert len(prm) == 1 and prm[0].gene_name == "test_panel_18S"
    assert prm[0] == PCRParams("TAACATATGCTTGTCTCAAAG", "CCTGTATTGTTATTTTTCGTCAC", gene_name="test_panel_18S", min_length=300,
                               max_length=700, min_count=2, mismatches=2, trim=15)
    with pytest.raises(panels.PanelError) as e:
        panels.load_panel_file("https://localhost:1/nonexistent_panel.yaml")
    assert "Failed to download panel from URL" in str(e.value)
    with pytest.raises(panels.PanelError) as e:
        panels.load_panel_file(str(tmp_path / "missing.yaml"))
    assert "Failed to read panel file" in str(e.value)


def test_schema_rules():
    v2 = """
name: no_clade_panel
schema_version: "2"
panel_version: "1.0.0"
description: "v2 panel without clade"
primers:
  - gene: "X"
    forward_seq: "AAAA"
    reverse_seq: "TTTT"
"""
    with pytest.raises(panels.PanelError) as e:
        panels.panel_to_params(panels.parse_panel_yaml(v2), "test")
    assert "clade" in str(e.value)
    assert panels.panel_to_params(panels.parse_panel_yaml(v2.replace('description:', 'clade: "Cnidaria"\ndescription:')), "t")[0].gene_name == "no_clade_panel_X"
    with pytest.raises(panels.PanelError) as e:
        panels.parse_panel_yaml('name: typo_panel\nversoin: 1.0.0\ndescription: "d"\nprimers:\n  - gene: "X"\n    forward_seq: "A"\n    reverse_seq: "T"\n')
    assert "unknown field `versoin`" in str(e.value)
    with pytest.raises(panels.PanelError) as e:
        panels.parse_panel_yaml('name: p\npanel_version: 1.0.0\ndescription: "d"\nprimers:\n  - gene: "X"\n    forward_seq: "A"\n    reverse_seq: "T"\n    forward_sqe: "oops"\n')
    assert "unknown field `forward_sqe`" in str(e.value)
    dup = 'name: dup_panel\npanel_version: "1.0.0"\ndescription: "d"\nprimers:\n  - gene: "CO1"\n    forward_seq: "AAAA"\n    reverse_seq: "TTTT"\n  - gene: "CO1"\n    forward_seq: "CCCC"\n    reverse_seq: "GGGG"\n'
    with pytest.raises(panels.PanelError) as e:
        panels.panel_to_params(panels.parse_panel_yaml(dup), "t")
    assert "duplicate" in str(e.value) and "positions 0 and 1" in str(e.value)


def test_names_prefix_deprecation_and_cli_specs(tmp_path):
    text = """
name: demo
schema_version: "2"
clade: "Cnidaria"
panel_version: "1.2.0"
gene_prefix: cn
description: "naming rules"
primers:
  - gene: "18S"
    region: "V4"
    index: 1
    forward_seq: "CCAGCASCYGCGGTAATTCC"
    reverse_seq: "ACTTTCGTTCTTGATYRA"
    max_length: 800
  - gene: "18S"
    region: "V4"
    index: 2
    forward_seq: "CCAGCASCYGCGGTAATTCC"
    reverse_seq: "ACTTTCGTTCTTGATYRR"
    deprecated: true
    deprecated_by: "cn_18S-V4_1"
    deprecated_reason: "superseded"
  - gene: "psbA-trnH"
    forward_seq: "GTTATGCATGAACGTAATGCTC"
    reverse_seq: "CGCGCATGGTGGATTCACAATCC"
    min_count: 3
    trim: 12
"""
    f = tmp_path / "demo.yaml"
    f.write_text(text)
    warnings = []
    prm = panels.load_panel_file(str(f), warn=warnings.append)
    assert [p.gene_name for p in prm] == ["cn_18S-V4_1", "cn_psbA-trnH"]
    assert warnings == ["Panel 'demo': skipping deprecated primer 'cn_18S-V4_2'. Use 'cn_18S-V4_1' instead. Reason: superseded"]
    assert prm[0].max_length == 800 and prm[0].min_count == 2 and prm[1].min_count == 3 and prm[1].trim == 12
    spec = panels.to_pcr_primers_spec(prm[1])
    assert spec == ("forward=GTTATGCATGAACGTAATGCTC,reverse=CGCGCATGGTGGATTCACAATCC,name=cn_psbA-trnH,min-length=0,"
                    "max-length=10000,min-count=3,mismatches=2,trim=12,dedup-edit-threshold=10")
    bad = text.replace('gene: "psbA-trnH"', 'gene: "psbA-trnH"\n    region: "x"')
    f.write_text(bad)
    with pytest.raises(panels.PanelError) as e:
        panels.load_panel_file(str(f))
    assert "Invalid primer specification in panel file" in str(e.value) and "must not contain '-' when a `region` is also set" in str(e.value)


def test_panel_drives_spcr(oracle, tmp_path):
    """A panel entry run through the pipeline: the FASTA is named after the derived, prefixed gene name."""
    import random
    from sharkmer_b200 import pcr
    from test_pcr import OracleTable, make_reads, rc
    rng = random.Random(4)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    fwd, rev, ins = rnd(20), rnd(20), rnd(250)
    genome = rnd(600) + fwd + ins + rc(rev) + rnd(600)
    t = oracle.KmerCounts(21)
    for s in make_reads(rng, genome, 900, 100):
        t.ingest_seq(s)
    f = tmp_path / "p.yaml"
    f.write_text(f'name: mini\ndescription: "d"\nprimers:\n  - gene: "CO1"\n    index: 3\n    forward_seq: "{fwd}"\n    reverse_seq: "{rev}"\n    max_length: 900\n')
    res = pcr.run_pcr(OracleTable(t), 21, panels.load_panel_file(str(f)), "smp", str(tmp_path) + "/")
    assert res[0]["gene_name"] == "mini_CO1_3" and res[0]["status"] == "success"
    assert open(tmp_path / "smp_mini_CO1_3.fasta").read().startswith(">smp_mini_CO1_3_0 sample=smp gene=mini_CO1_3 product=0 ")


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/panels"), reason="reference checkout not mounted")
def test_reference_panels_load_unchanged():
    """The reference's own panel files (data, read in place, never copied) load with these rules:
    every panel parses, names are unique, and the cnidaria panel of BASELINE config C4 is among them."""
    import glob
    import os
    seen = {}
    for path in sorted(glob.glob("/root/reference/panels/*.yaml")):
        prm = panels.load_panel_file(path)
        assert prm, path
        names = [p.gene_name for p in prm]
        assert len(set(names)) == len(names)
        for p in prm:
            from sharkmer_b200 import pcr
            assert pcr.validate_pcr_params(p) == [], (path, p.gene_name)
        seen[os.path.basename(path)] = len(prm)
    assert "cnidaria.yaml" in seen and seen["cnidaria.yaml"] >= 5


@pytest.mark.skipif(not __import__("os").path.isfile("/root/reference/panels/cnidaria.yaml"), reason="reference checkout not mounted")
def test_c4_cnidaria_panel_on_planted_templates(oracle, tmp_path):
    """BASELINE config C4 in miniature (k = 25, the cnidaria panel): amplicon templates built from each
    primer pair (one concrete reading of the degenerate primers, a random insert inside the pair's
    [min_length, max_length]) are planted in a random genome, reads are sampled with errors, counted,
    and every gene of the panel must come back base for base."""
    import random
    from sharkmer_b200 import pcr
    from sharkmer_b200.primers import _IUPAC
    from test_pcr import OracleTable, make_reads, rc
    k = 25
    rng = random.Random(25)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    concrete = lambda s: "".join("ACGT"[rng.choice(_IUPAC[c])] for c in s)
    prm = panels.load_panel_file("/root/reference/panels/cnidaria.yaml")
    genome, truth = rnd(2000), {}
    for p in prm:
        f, r = concrete(p.forward_seq), concrete(p.reverse_seq)
        lo = max(p.min_length, len(f) + len(r) + 60)
        total = min(max(lo, min(p.max_length, lo + 300)), p.max_length)
        insert = rnd(max(40, total - len(f) - len(r)))
        amplicon = f + insert + rc(r)
        trim_f, trim_r = min(p.trim, k - 1, len(f)), min(p.trim, k - 1, len(r))
        truth[p.gene_name] = amplicon[len(f) - trim_f:len(amplicon) - (len(r) - trim_r)]
        genome += amplicon + rnd(1500)
    t = oracle.KmerCounts(k)
    n_reads = len(genome) * 30 // 120
    for s in make_reads(rng, genome, n_reads, 120, err=0.002):
        t.ingest_seq(s)
    res = pcr.run_pcr(OracleTable(t), k, prm, "c4", str(tmp_path) + "/")
    got = {r["gene_name"]: r for r in res}
    assert set(got) == set(truth)
    for name, want in truth.items():
        assert got[name]["status"] == "success", (name, got[name])
        fa = open(tmp_path / f"c4_{name}.fasta").read().split("\n")
        first = []
        for line in fa[1:]:
            if line.startswith(">") or not line:
                break
            first.append(line)
        assert "".join(first) == want, name
