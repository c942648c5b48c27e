"""Extracts the seven RemoveSoftLineBreaks(32) cases of
/root/reference/packages/circuits/tests/remove-soft-line-breaks.test.ts into tests/golden/remove_soft_line_breaks.json.
Run in the build container (the reference tree is not present on the GPU box)."""
import json, re, sys
src = open("/root/reference/packages/circuits/tests/remove-soft-line-breaks.test.ts").read()
src = re.sub(r"//[^\n]*", "", src)   # drop line comments
cases = []
for m in re.finditer(r"it\('([^']+)'.*?const input = \{(.*?)\};.*?isValid: (\d)", src, re.S):
    name, body, valid = m.group(1), m.group(2), int(m.group(3))
    def arr(key):
        a = re.search(key + r":\s*\[(.*?)\]\s*,?\s*(?:decoded|\Z|\})", body + "}", re.S).group(1)
        out = []
        for tok in re.split(r",", a):
            tok = tok.strip()
            if not tok:
                continue
            f = re.match(r"\.\.\.Array\((\d+)\)\.fill\((\d+)\)", tok)
            if f:
                out += [int(f.group(2))] * int(f.group(1))
            else:
                out.append(int(tok))
        return out
    enc = arr("encoded")
    dec_src = body[body.index("decoded"):]
    a = re.search(r"decoded:\s*\[(.*?)\]", dec_src, re.S).group(1)
    dec = []
    for tok in a.split(","):
        tok = tok.strip()
        if not tok:
            continue
        f = re.match(r"\.\.\.Array\((\d+)\)\.fill\((\d+)\)", tok)
        dec += [int(f.group(2))] * int(f.group(1)) if f else [int(tok)]
    cases.append({"name": name, "encoded": enc, "decoded": dec, "isValid": valid})
json.dump(cases, open(sys.argv[1], "w"), indent=1)
print(len(cases), [(c["name"], len(c["encoded"]), len(c["decoded"]), c["isValid"]) for c in cases])
