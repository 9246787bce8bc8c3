#!/usr/bin/env python3
"""stdin -> stdout filter used by oracle/build_ref.sh for the *_pade flavours.

Turns the fork's step_generic.cpp (or its generated stride-1 twin) back into upstream Meep's
Kerr/Pockels behaviour by RESTORING THE FORMULAS THE FORK COMMENTED OUT -- nothing is invented:
  * calc_nonlinear_u: drop the `cout << "shuoldn't be here pade!"` / `sleep(10)` lines (:550-551)
  * `if (false) { //if (chi3) {`  ->  `if (chi3) {`                                   (:608,638,668)
  * every `f[i] = (gs * us);` / `fw[i] = (gs * us);` that is directly followed by a commented-out
    alternative `// f[i] = ...;` is replaced by that alternative   (:632,646-647,659,823,835-836,
    844,862-863,873-874,884)
  * the Newton-Raphson loop body of the non-PML 3x3 branch (:736-813) is replaced by the upstream
    formula kept in the block comment at :768-773
TEST INFRASTRUCTURE: its output only ever exists in the compiler's input pipe.
"""
import re
import sys

src = sys.stdin.read()

# 1. calc_nonlinear_u
src = re.sub(r'\n\s*cout << "shuoldn\'t be here pade!" <<endl;\n\s*sleep\(10\);', "", src)

# 2. disabled chi3 tests in the PML branches
src = re.sub(r"if \(false\) \{\s*//\s*if \(chi3\) \{[^\n]*", "if (chi3) {", src)

# 4. NR loop body -> upstream formula (do this before step 3 so that the block comment survives)
start = src.index("// NEW STUFF Here")
end = src.index("// TODO REPLACE STuFF HERE ^^^^^^^^^^")
blk = src[start:end]
m = re.search(r"/\*\s*(realnum g1s = .*?chi3\[i\]\);)\s*\*/", blk, re.S)
assert m, "upstream formula comment not found"
formula = m.group(1)
# the region [start, end) sits inside `PLOOP... {` ... and is followed by `}` closing the else{} of
# the NR test; replace the region and that one brace
tail = src[end:]
tail = tail[tail.index("\n") + 1:]           # rest of the TODO line
tail = tail[tail.index("}") + 1:]            # closes `else {` of the NR branch
src = src[:start] + "\n          " + formula + "\n" + tail

# 3. restore commented-out alternatives
lines = src.split("\n")
out = []
i = 0
pat_active = re.compile(r"^(\s*)(fw|f)\[i\] = \(gs \* us\);\s*$")
while i < len(lines):
    m = pat_active.match(lines[i])
    if m and i + 1 < len(lines):
        j = i + 1
        com = []
        while j < len(lines) and re.match(r"^\s*//", lines[j]) and len(com) < 3:
            com.append(re.sub(r"^\s*//\s?", "", lines[j]))
            j += 1
        text = " ".join(com)
        k = text.find(m.group(2) + "[i] =")
        if com and k >= 0 and ";" in text[k:]:
            stmt = text[k:text.rindex(";") + 1]
            # cut trailing chatter after the statement's own terminating ';'
            depth, endpos = 0, None
            for pos, ch in enumerate(stmt):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == ";" and depth == 0:
                    endpos = pos
                    break
            stmt = stmt[:endpos + 1]
            out.append(m.group(1) + stmt)
            i = j
            continue
    out.append(lines[i])
    i += 1
sys.stdout.write("\n".join(out))
