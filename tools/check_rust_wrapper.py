#!/usr/bin/env python
"""Static agreement checks of the hand-written Rust wrapper (rust/rusterix-cuda/src/{cuda,lower}.rs) with the generated
-sys crate and the header -- what can be verified without a Rust toolchain:
  * every rxc_* function the wrapper calls is declared in rusterix-cuda-sys with that many arguments;
  * every rxc_* struct literal names exactly the fields of the mirrored struct (no field missing, none unknown);
  * every RXC_* / RXVM_* constant it uses exists in the -sys crate;
  * lower.rs's NodeOp -> opcode table equals rusterix_b200/vm.py's OPS (= the header's RXVM_* numbering) and, when the
    reference checkout is present, the declaration order of `enum NodeOp` (rusteria/src/node/nodeop.rs);
  * the flat-code opcodes OP_JZ .. OP_END equal RXVM_JZ .. RXVM_END.
Prints the findings and exits 1 on any mismatch."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rusterix_b200 import vm  # noqa: E402


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def matching(text, start, open_ch, close_ch):
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced")


def split_top(body):
    parts, depth, cur = [], 0, ""
    for ch in body:
        if ch in "([{":          # (angle brackets are not tracked: `=>` and `->` would unbalance them, and no generic
            depth += 1           #  argument list of these files holds a top-level comma)
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


def main():
    sys_rs = open(os.path.join(ROOT, "rust", "rusterix-cuda-sys", "src", "lib.rs")).read()
    fns = {m.group(1): len(split_top(m.group(2))) for m in re.finditer(r"pub fn (rxc_\w+)\((.*?)\)", sys_rs, flags=re.S)}
    structs = {m.group(1): re.findall(r"pub (\w+):", m.group(2)) for m in re.finditer(r"pub struct (rxc_\w+) \{(.*?)\n\}", sys_rs, flags=re.S)}
    consts = set(re.findall(r"pub const (\w+):", sys_rs))
    errors, checked = [], {"calls": 0, "literals": 0, "constants": 0}
    for name in ("cuda.rs", "lower.rs"):
        src = strip_comments(open(os.path.join(ROOT, "rust", "rusterix-cuda", "src", name)).read())
        src = re.sub(r'"(?:[^"\\]|\\.)*"', '""', src)            # string literals hold braces and commas
        for m in re.finditer(r"\b(rxc_\w+)\s*\(", src):
            fn = m.group(1)
            if fn in structs:
                continue
            end = matching(src, m.end() - 1, "(", ")")
            n = len(split_top(src[m.end():end]))
            checked["calls"] += 1
            if fn not in fns:
                errors.append(f"{name}: calls {fn}, which the -sys crate does not declare")
            elif fns[fn] != n:
                errors.append(f"{name}: {fn} called with {n} arguments, declared with {fns[fn]}")
        for m in re.finditer(r"\b(rxc_\w+)\s*\{", src):
            st = m.group(1)
            if st not in structs:
                continue
            before = src[max(0, m.start() - 12):m.start()]
            if re.search(r"(->|:|&|<|\*const|\*mut)\s*$", before):     # a type position, not a literal
                continue
            end = matching(src, m.end() - 1, "{", "}")
            fields = [re.match(r"(\w+)", p).group(1) for p in split_top(src[m.end():end])]
            checked["literals"] += 1
            if sorted(fields) != sorted(structs[st]):
                missing = sorted(set(structs[st]) - set(fields)); extra = sorted(set(fields) - set(structs[st]))
                errors.append(f"{name}: literal of {st}: missing {missing}, unknown {extra}")
        for c in set(re.findall(r"\b(RX(?:C|VM)_[A-Z0-9_]+)\b", src)):
            checked["constants"] += 1
            if c not in consts:
                errors.append(f"{name}: uses {c}, which the -sys crate does not define")
    lower = strip_comments(open(os.path.join(ROOT, "rust", "rusterix-cuda", "src", "lower.rs")).read())
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b([A-Z]\w*)(?:\([^)]*\))?\s*=>\s*(\d+)\s*[,}]", lower[lower.index("pub fn opcode"):lower.index("fn swizzle_get")])}
    want = {n: i for i, n in enumerate(vm.OPS[:90])}
    if table != want:
        errors.append(f"lower.rs: opcode table differs from vm.OPS: {sorted(set(table.items()) ^ set(want.items()))[:6]}")
    for k, nm in enumerate(("JZ", "JMP", "MARK", "TRUNC", "UNMARK", "END")):
        m = re.search(r"pub const OP_%s: u32 = (\d+);" % nm, lower)
        if not m or int(m.group(1)) != 90 + k or vm.OPCODE[{"JZ": "Jz", "JMP": "Jmp", "MARK": "Mark", "TRUNC": "Trunc", "UNMARK": "Unmark", "END": "End"}[nm]] != 90 + k:
            errors.append(f"lower.rs: OP_{nm} is not {90 + k}")
    ref = "/root/reference/rusteria/src/node/nodeop.rs"
    if os.path.exists(ref):
        body = strip_comments(open(ref).read())
        body = body[body.index("pub enum NodeOp"):]
        body = body[body.index("{") + 1:matching(body, body.index("{"), "{", "}")]
        order = [re.match(r"(\w+)", p).group(1) for p in split_top(body)]
        if order != vm.OPS[:90]:
            errors.append("reference NodeOp declaration order differs from vm.OPS")
        checked["reference_nodeop_variants"] = len(order)
    print("checked", checked)
    for e in errors:
        print("MISMATCH:", e)
    return 1 if errors else 0


if __name__ == "__main__":
    sys.exit(main())
