"""Primer panels for sPCR: user-supplied panel YAML files -> PCRParams, with the reference's schema
checks and output-name rules (caseywdunn/sharkmer v3.1.0, src/pcr/preconfigured.rs):

    deny_unknown_fields on the panel and on every primer      :6-70, pcr/mod.rs:148-150
    schema_version "2" requires `clade`                        :322-334
    gene must not contain '_', nor '-' when a region is set; region must not contain '_'   :181-213
    (gene, region, index) unique within a panel                :216-240
    gene_name = {gene}[-{region}][_{index}], prefixed with gene_prefix or the panel name   :170-178, 289-292
    deprecated primers are skipped                             :339-360

The reference also embeds nine built-in panels (`--pcr-panel cnidaria`, :427-446); those are data
files of the reference repository and are not redistributed here — point `load_panel_file` at a
panel file (the reference's own panels load unchanged).  Panels over HTTP (:371-420) are not
supported: no network.

`python -m sharkmer_b200.panels panel.yaml` prints one `--pcr-primers "..."` argument per primer
pair for the C++ driver (sharkmer_b200_cli), which has no YAML parser.
"""
from __future__ import annotations

import sys

from .primers import PCRParams

PANEL_FIELDS = {"name", "schema_version", "panel_version", "description", "clade", "taxon_id", "gene_prefix", "status",
                "source_url", "license", "citation", "notes", "$schema", "maintainers", "changelog", "primers",
                "references", "validation"}
PRIMER_FIELDS = {"forward_seq", "reverse_seq", "min_length", "max_length", "gene", "region", "index", "compartment",
                 "gene_type", "copy_number", "deprecated", "deprecated_by", "deprecated_reason", "min_count",
                 "mismatches", "trim", "expected_length", "citation", "notes", "dedup_edit_threshold",
                 "max_dfs_states", "max_paths_per_pair", "max_node_visits", "max_primer_kmers", "high_coverage_ratio",
                 "tip_coverage_fraction"}
_PARAM_FIELDS = ("forward_seq", "reverse_seq", "min_length", "max_length", "min_count", "mismatches", "trim",
                 "dedup_edit_threshold", "max_dfs_states", "max_paths_per_pair", "max_node_visits", "max_primer_kmers",
                 "high_coverage_ratio", "tip_coverage_fraction")


class PanelError(ValueError):
    pass


def is_url(source: str) -> bool:   # :243-245
    return source.startswith("http://") or source.startswith("https://")


def derive_gene_name(gene: str, region=None, index=None) -> str:   # :170-178
    name = gene
    if region is not None:
        name += f"-{region}"
    if index is not None:
        name += f"_{index}"
    return name


def validate_gene_chars(gene: str, has_region: bool):   # :186-203
    if "_" in gene:
        raise PanelError(f"gene '{gene}' must not contain '_' (reserved as index delimiter in output names).")
    if has_region and "-" in gene:
        raise PanelError(
            f"gene '{gene}' must not contain '-' when a `region` is also set, because the derived output name "
            "`{gene}-{region}` would be ambiguous. Use alphanumeric characters only for `gene` when pairing it with a "
            "`region` (e.g. 'CytB' not 'Cyt-b').")


def validate_region_chars(region: str):   # :206-213
    if "_" in region:
        raise PanelError(f"region '{region}' must not contain '_' (reserved as index delimiter in output names).")


def parse_panel_yaml(text: str) -> dict:   # :165-167 + the deny_unknown_fields of the structs
    import yaml
    try:
        panel = yaml.safe_load(text)
    except yaml.YAMLError as e:
        raise PanelError(f"Failed to parse panel YAML: {e}") from e
    if not isinstance(panel, dict):
        raise PanelError("Failed to parse panel YAML: expected a mapping")
    for key in panel:
        if key not in PANEL_FIELDS:
            raise PanelError(f"Failed to parse panel YAML: unknown field `{key}`")
    for key in ("name", "description", "primers"):
        if key not in panel:
            raise PanelError(f"Failed to parse panel YAML: missing field `{key}`")
    if not isinstance(panel["primers"], list):
        raise PanelError("Failed to parse panel YAML: `primers` must be a sequence")
    for p in panel["primers"]:
        if not isinstance(p, dict):
            raise PanelError("Failed to parse panel YAML: a primer must be a mapping")
        for key in p:
            if key not in PRIMER_FIELDS:
                raise PanelError(f"Failed to parse panel YAML: unknown field `{key}`")
        for key in ("forward_seq", "reverse_seq"):
            if key not in p:
                raise PanelError(f"Failed to parse panel YAML: missing field `{key}`")
    return panel


def require_clade_for_v2(panel: dict, source: str):   # :322-334
    if str(panel.get("schema_version")) == "2" and panel.get("clade") is None:
        raise PanelError(
            f"Panel '{panel['name']}' from {source} declares schema_version: \"2\" but is missing the required `clade` "
            "field. Set `clade` to the NCBI-preferred taxon name for the target clade (e.g. `clade: \"Cnidaria\"`).")


def resolve_primer_gene_names(primers: list, panel_name: str):   # :216-278 -> list of gene names ('' when no `gene`)
    for p in primers:
        if p.get("gene") is not None:
            validate_gene_chars(str(p["gene"]), p.get("region") is not None)
        if p.get("region") is not None:
            validate_region_chars(str(p["region"]))
    seen = {}
    for i, p in enumerate(primers):
        if p.get("gene") is None:
            continue
        key = (str(p["gene"]), None if p.get("region") is None else str(p["region"]), p.get("index"))
        if key in seen:
            def q(v):
                return "None" if v is None else (f'Some("{v}")' if isinstance(v, str) else f"Some({v})")
            raise PanelError(
                f"Panel '{panel_name}': duplicate primer entries for (gene=\"{key[0]}\", region={q(key[1])}, "
                f"index={q(key[2])}) at positions {seen[key]} and {i}. Add an `index:` field to distinguish them.")
        seen[key] = i
    return [derive_gene_name(str(p["gene"]), None if p.get("region") is None else str(p["region"]), p.get("index"))
            if p.get("gene") is not None else "" for p in primers]


def panel_to_params(panel: dict, source: str, warn=None):
    """The primers of a parsed panel as PCRParams, named and filtered as load_panel_file does (:248-266)."""
    require_clade_for_v2(panel, source)
    names = resolve_primer_gene_names(panel["primers"], panel["name"])
    prefix = panel.get("gene_prefix") or panel["name"]
    out = []
    for p, name in zip(panel["primers"], names):
        full = f"{prefix}_{name}"
        if p.get("deprecated"):
            if warn:
                msg = f"Panel '{panel['name']}': skipping deprecated primer '{full}'."
                if p.get("deprecated_by"):
                    msg += f" Use '{p['deprecated_by']}' instead."
                if p.get("deprecated_reason"):
                    msg += f" Reason: {p['deprecated_reason']}"
                warn(msg)
            continue
        kw = {k: p[k] for k in _PARAM_FIELDS if k in p}
        out.append(PCRParams(gene_name=full, **kw))
    return out


def load_panel_file(path: str, warn=None):   # :248-266
    if is_url(path):
        raise PanelError(f"Failed to download panel from URL: {path} (network error, timeout, or HTTP failure)")
    try:
        text = open(path).read()
    except OSError as e:
        raise PanelError(f"Failed to read panel file: {path}") from e
    try:
        panel = parse_panel_yaml(text)
    except PanelError as e:
        raise PanelError(f"Failed to parse panel file '{path}'. Check for YAML syntax errors and ensure all primer "
                         f"fields are valid. ({e})") from e
    try:
        return panel_to_params(panel, path, warn)
    except PanelError as e:
        if "schema_version" in str(e):
            raise
        raise PanelError(f"Invalid primer specification in panel file '{path}': {e}") from e


def to_pcr_primers_spec(p: PCRParams) -> str:
    """The `--pcr-primers` argument (src/cli.rs:12-140) for one primer pair."""
    return (f"forward={p.forward_seq},reverse={p.reverse_seq},name={p.gene_name},min-length={p.min_length},"
            f"max-length={p.max_length},min-count={p.min_count},mismatches={p.mismatches},trim={p.trim},"
            f"dedup-edit-threshold={p.dedup_edit_threshold}")


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        for prm in load_panel_file(arg, warn=lambda m: print(m, file=sys.stderr)):
            print(f'--pcr-primers "{to_pcr_primers_spec(prm)}"')
