#!/usr/bin/env python
"""tests/golden/cli_flags.json: the command-line flags of the reference scripts whose entry points this package mirrors
(names, defaults, types, required / nargs / action), read STATICALLY from the reference's sources with `ast` — nothing
of the reference is imported or executed.  tests/test_host_logic.py::test_cli_flags_match_the_reference compares the
package's own parsers with it.

    python tests/golden/make_cli_flags.py            (needs /root/reference; run in the build container)
"""
import ast
import json
import os

REF = os.environ.get("SFR_REFERENCE_ROOT", "/root/reference")
SCRIPTS = {
    "dit_forget": "DiT/forget.py",
    "dit_generate_fisher": "DiT/generate_fisher.py",
    "dit_generate_mask": "DiT/generate_mask.py",
    "sd_generate_fisher": "SD/train-scripts/generate_fisher.py",
    "sd_nsfw_removal": "SD/train-scripts/nsfw_removal.py",
    "sd_generate_fisher_mask": "SD/train-scripts/generate_fisher_mask.py",
    "ddpm_generate_fisher_mask": "DDPM/generate_fisher_mask.py",
}


def literal(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return "<expr> " + ast.unparse(node)


def flags_of(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument"):
            continue
        names = [a.value for a in node.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
        if not names:
            continue
        rec = {"flags": names, "line": node.lineno}
        for kw in node.keywords:
            if kw.arg == "help":
                continue
            if kw.arg == "type":
                rec["type"] = ast.unparse(kw.value)
            else:
                rec[kw.arg] = literal(kw.value)
        out.append(rec)
    out.sort(key=lambda r: r["line"])
    return out


def main():
    doc = {name: {"source": rel, "flags": flags_of(os.path.join(REF, rel))} for name, rel in SCRIPTS.items()}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_flags.json")
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
        f.write("\n")
    for name, d in doc.items():
        print(name, len(d["flags"]), "flags")


if __name__ == "__main__":
    main()
