#!/usr/bin/env python
"""abi_index.py — the index of include/t4k.h printed as the markdown table of INTEGRATION.md §8: every exported function, the
section of the header it stands in (each section names the reference files it replaces) and the reference lines its own
comment cites.  `python bench_scripts/abi_index.py --write` regenerates the block between the two markers in INTEGRATION.md;
tests/test_abi.py keeps the table and the header in step."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BEGIN, END = "<!-- abi-index:begin -->", "<!-- abi-index:end -->"
DECL = re.compile(r"^(?:const char\s*\*\s*|(?:int64_t|int|long)\s+)(t4k_[a-z0-9_]+)\s*\(", re.M)
CITE = re.compile(r"(?:src/|examples/)?[A-Za-z0-9_/]+\.(?:cu|cpp|h|tcu|4th):\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*")
SECT = re.compile(r"^/\* ---- (.*?)(?: -+)?\s*(?:\*/)?$", re.M)


def entries():
    src = open(os.path.join(ROOT, "include", "t4k.h")).read()
    sects = [(m.start(), m.group(1).strip(" -")) for m in SECT.finditer(src)]
    decls = [(m.start(), m.group(1)) for m in DECL.finditer(src)]
    out = []
    prev_end = 0
    for i, (pos, name) in enumerate(decls):
        sec = [s for s in sects if s[0] < pos]
        sec = sec[-1][1] if sec else ""
        nxt = decls[i + 1][0] if i + 1 < len(decls) else len(src)
        end = src.find(";", pos)
        line_end = src.find("\n", end)
        if "/*" in src[end:line_end] and "*/" not in src[end:line_end]:      # a trailing comment that runs over several lines
            line_end = min(src.find("*/", end) + 2, nxt)
        # context = the comment in front of the declaration (back to the previous declaration or the section header) + the trailing comment
        sec_pos = max([s[0] for s in sects if s[0] < pos] or [0])
        ctx = src[max(prev_end, sec_pos):line_end]
        prev_end0, prev_end = prev_end, line_end
        cites = []
        for c in CITE.findall(ctx):
            if c not in cites:
                cites.append(c)
        if not cites and "/*" not in src[max(prev_end0, sec_pos):pos] and out and out[-1][1] == sec:
            cites = out[-1][2]                                             # declared under the previous declaration's comment (a group)
        if not cites:
            cites = []                                                     # the section header's own citations (its whole comment)
            for c in CITE.findall(src[sec_pos:src.find("*/", sec_pos)]):
                if c not in cites and not c.startswith("tensorforth_b200"):
                    cites.append(c)
        out.append((name, sec, cites))
    return out


def table():
    rows = ["| entry point | header section (reference files it replaces) | reference lines cited at the declaration |", "|---|---|---|"]
    for name, sec, cites in entries():
        sec = re.sub(r"\s+", " ", sec)
        if len(sec) > 110:
            sec = sec[:107] + "..."
        rows.append("| `%s` | %s | %s |" % (name, sec, ", ".join("`%s`" % c for c in cites[:6]) or "— (new: the reference has no counterpart)"))
    return "\n".join(rows)


if __name__ == "__main__":
    t = table()
    if "--write" in sys.argv:
        p = os.path.join(ROOT, "INTEGRATION.md")
        s = open(p).read()
        a, b = s.index(BEGIN) + len(BEGIN), s.index(END)
        open(p, "w").write(s[:a] + "\n" + t + "\n" + s[b:])
    else:
        print(t)
