"""Single-variable freedom probe of an R1CS at a valid witness (test infrastructure).

An independent look at the FRONT END (csrc/gadgets.cpp, csrc/regex.cpp): it uses nothing but the exported constraint
matrices and one satisfying witness - not the witness program the GPU interpreter and oracle/zkref_witness.c both walk.
For every variable v it asks whether the constraint system admits another value for v with all other variables fixed:
with w' = w + delta e_v, row i becomes  delta (a_iv <B_i,w> + b_iv <A_i,w> - c_iv) + delta^2 a_iv b_iv = 0, so v is
  * FREE        if g_i = a_iv <B_i,w> + b_iv <A_i,w> - c_iv and q_i = a_iv b_iv vanish in every row that mentions v,
  * TWO-VALUED  if one delta != 0 solves g_i + delta q_i = 0 in every such row.
A gadget that forgets to constrain a hinted signal (the classic circom bug: `<--` without `===`) shows up as FREE.
Some freedom is legitimate and the tests list it by class: the inverse hint of IsZero at in = 0, a select of equal
values, products with a factor that is zero at this witness.
"""
from __future__ import annotations

from zkemail_b200 import FR_MODULUS as R
from zkemail_b200 import _lib as L
from zkemail_b200.iden3_binfile import _coefs, _u32_array


def matrices(circuit):
    coefs = [int.from_bytes(c, "little") for c in _coefs(circuit)]
    out = []
    for p, v, c in ((L.ARR_A_PTR, L.ARR_A_VAR, L.ARR_A_COEF), (L.ARR_B_PTR, L.ARR_B_VAR, L.ARR_B_COEF),
                    (L.ARR_C_PTR, L.ARR_C_VAR, L.ARR_C_COEF)):
        out.append((list(_u32_array(circuit, p)), list(_u32_array(circuit, v)), [coefs[k] for k in _u32_array(circuit, c)]))
    return out


def probe(circuit, w, drop_rows=()):
    """w: list of n_vars integers (a satisfying witness).  Returns (free, two_valued, unmentioned): free / two_valued
    map a variable id to the first constraint row that mentions it, unmentioned is a list of variable ids.
    drop_rows: constraint rows to leave out (the probe's own test removes one to see the probe object)."""
    n_rows, n_vars = circuit.info.n_constraints, circuit.info.n_vars
    (ap, av, ac), (bp, bv, bc), (cp, cv, cc) = matrices(circuit)
    rows_of = [None] * n_vars          # v -> {row: [a, b, c]}
    aw, bw = [0] * n_rows, [0] * n_rows
    for which, (ptr, var, coef) in enumerate(((ap, av, ac), (bp, bv, bc), (cp, cv, cc))):
        for i in range(n_rows):
            if i in drop_rows:
                continue
            s = 0
            for k in range(ptr[i], ptr[i + 1]):
                v, c = var[k], coef[k]
                s += c * w[v]
                d = rows_of[v]
                if d is None:
                    d = rows_of[v] = {}
                e = d.get(i)
                if e is None:
                    e = d[i] = [0, 0, 0]
                e[which] = (e[which] + c) % R
            if which == 0:
                aw[i] = s % R
            elif which == 1:
                bw[i] = s % R
    free, two, unmentioned = {}, {}, []
    n_out = circuit.info.n_outputs
    first_input, n_inputs = 1 + n_out, circuit.info.n_pub_inputs + circuit.info.n_prv_inputs
    for v in range(1, n_vars):
        if first_input <= v < first_input + n_inputs:
            continue                   # inputs are the statement's free variables by definition
        d = rows_of[v]
        if d is None:
            unmentioned.append(v)
            continue
        delta, ok, all_zero = None, True, True
        for i, (a, b, c) in d.items():
            g = (a * bw[i] + b * aw[i] - c) % R
            q = a * b % R
            if g == 0 and q == 0:
                continue
            all_zero = False
            if q == 0:
                ok = False
                break
            dl = (-g) * pow(q, -1, R) % R
            if dl == 0 or (delta is not None and dl != delta):
                ok = False
                break
            delta = dl
        if all_zero:
            free[v] = min(d)
        elif ok and delta is not None:
            two[v] = min(d)
    return free, two, unmentioned


def scopes_of(circuit, var_to_row):
    """{scope name: count} for a probe result (scope = the gadget instance that emitted the variable's first row)."""
    import ctypes
    p, n = circuit.array(L.ARR_SCOPE_OF_CONSTRAINT, None)
    scopes = (ctypes.c_uint16 * n).from_address(p)
    out = {}
    for v, row in var_to_row.items():
        name = circuit.scope_name(scopes[row])
        out[name] = out.get(name, 0) + 1
    return out
