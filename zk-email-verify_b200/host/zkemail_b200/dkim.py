"""Offline DKIM (RFC 6376) verification producing the reference's `DKIMVerificationResult`.

Mirrors `verifyDKIMSignature` (/root/reference/packages/helpers/src/dkim/index.ts:36-97) and the parts of the
modified mailauth it drives: message splitting (lib/mailauth/message-parser.ts:43-113), signed-header selection
bottom-up per h= (lib/mailauth/tools.ts:107-140), relaxed header canonicalisation (tools.ts:441-454,
header/relaxed.ts:5-81 incl. emptying b=), relaxed / simple body canonicalisation (body/relaxed.ts:157-273,
body/simple.ts:59-106), plus the sanitizer retry loop (dkim/index.ts:49-66 with dkim/sanitizers.ts:1-67: when the first
attempt ends in "bad signature" every sanitizer is tried and the first passing one is reported as
`appliedSanitization`).  Difference by design: key lookup goes through a caller-supplied `resolver` (the reference does
DNS-over-HTTPS with an archive fallback, dkim/dns-over-http.ts:100-156, dkim/index.ts:105-131 - no network here); all
records the resolver returns are tried in turn.  CPU-side I/O and string work; not part of the GPU hot path.
"""
from __future__ import annotations
import base64
import hashlib
import re
from dataclasses import dataclass

from cryptography.hazmat.primitives import hashes, serialization
from cryptography.hazmat.primitives.asymmetric import padding, rsa
from cryptography.exceptions import InvalidSignature


@dataclass
class DKIMVerificationResult:  # dkim/index.ts:12-24
    publicKey: int
    signature: int
    headers: bytes
    body: bytes
    bodyHash: str
    signingDomain: str
    selector: str
    algo: str
    format: str
    modulusLength: int
    appliedSanitization: str | None = None


def _normalise_newlines(raw: bytes) -> bytes:
    # message-parser.ts: bare LF -> CRLF
    return re.sub(rb"(?<!\r)\n", b"\r\n", raw)


def split_message(raw: bytes):
    raw = _normalise_newlines(raw)
    idx = raw.find(b"\r\n\r\n")
    if idx < 0:
        head, body = raw, b""
    else:
        head, body = raw[:idx], raw[idx + 4:]
    lines = head.split(b"\r\n")
    headers = []
    for ln in lines:
        if ln[:1] in (b" ", b"\t") and headers:
            headers[-1] = headers[-1] + b"\r\n" + ln
        elif ln:
            headers.append(ln)
    parsed = []
    for h in headers:
        key = h.split(b":", 1)[0].strip().lower().decode("latin-1")
        parsed.append((key, h))
    return parsed, body


def format_relaxed_line(line: bytes, suffix: bytes = b"") -> bytes:
    """formatRelaxedLine (tools.ts:441-454)."""
    s = line.decode("latin-1")
    s = re.sub(r"\r?\n", "", s)
    s = re.sub(r"^([^:]*):\s*", lambda m: m.group(1).lower().strip() + ":", s, count=1)
    s = re.sub(r"\s+", " ", s).strip()
    return s.encode("latin-1") + suffix


def relaxed_body(body: bytes) -> bytes:
    """RFC 6376 3.4.4 (body/relaxed.ts): strip trailing WSP per line, collapse WSP runs, drop trailing empty lines,
    terminate a non-empty body with CRLF."""
    lines = body.split(b"\r\n")
    out = []
    for ln in lines:
        ln = re.sub(rb"[ \t]+", b" ", ln)
        ln = re.sub(rb" +$", b"", ln)
        out.append(ln)
    while out and out[-1] == b"":
        out.pop()
    if not out:
        return b""
    return b"\r\n".join(out) + b"\r\n"


def simple_body(body: bytes) -> bytes:
    """RFC 6376 3.4.3 (body/simple.ts): drop trailing empty lines; empty body -> CRLF."""
    while body.endswith(b"\r\n\r\n"):
        body = body[:-2]
    if not body:
        return b"\r\n"
    if not body.endswith(b"\r\n"):
        body += b"\r\n"
    return body


def parse_tag_list(value: str) -> dict:
    tags = {}
    for part in value.split(";"):
        if "=" not in part:
            continue
        k, v = part.split("=", 1)
        tags[k.strip()] = v.strip()
    return tags


def signed_header_bytes(parsed, dkim_line: bytes, h_list: str, header_canon: str) -> bytes:
    names = [k.strip().lower() for k in h_list.split(":") if k.strip()]
    pool = list(parsed)
    chunks = []
    for name in names:  # tools.ts:115-127 - pick bottom-up, each instance used once
        for i in range(len(pool) - 1, -1, -1):
            if pool[i][0] == name:
                line = pool[i][1]
                chunks.append(format_relaxed_line(line, b"\r\n") if header_canon == "relaxed" else line + b"\r\n")
                pool.pop(i)
                break
    if header_canon == "relaxed":
        sig = format_relaxed_line(dkim_line).decode("latin-1")
        sig = re.sub(r"([;:\s]+b=)[^;]+", r"\1", sig, count=1)   # header/relaxed.ts:70-78
        chunks.append(sig.encode("latin-1"))
    else:
        sig = dkim_line.decode("latin-1")
        sig = re.sub(r"([;:\s]+b=)[^;]+", r"\1", sig, count=1)
        chunks.append(sig.encode("latin-1"))
    return b"".join(chunks)


def _offline_resolver(name, rtype):
    raise LookupError("No DNS records found from any source")


# ---- dkim/sanitizers.ts:1-67 -------------------------------------------------------------------------------------
def _get_header_value(email: str, header: str) -> str:
    start = email.find(f"{header}: ")
    if start < 0:
        return ""
    start += len(header) + 2
    end = email.find("\n", start)
    return email[start:end if end >= 0 else len(email)]


def _set_header_value(email: str, header: str, value: str) -> str:
    old = _get_header_value(email, header)
    return email.replace(old, value, 1) if old else email


def revertGoogleMessageId(email: str) -> str:   # sanitizers.ts:16-29
    """Google replaces Message-ID when it ARC-forwards and keeps the original in X-Google-Original-Message-ID."""
    if "ARC-Authentication-Results" not in email:
        return email
    original = _get_header_value(email, "X-Google-Original-Message-ID")
    if original:
        return _set_header_value(email, "Message-ID", original)
    return email


def removeLabels(email: str) -> str:            # sanitizers.ts:32-36
    """`Subject: [ListName] Newsletter` -> `Subject: Newsletter` (the JS regex is greedy, so is this one)."""
    return re.sub(r"Subject: \[.*\]", "Subject:", email, count=1)


def insert13Before10(email: str) -> str:        # sanitizers.ts:40-57
    return re.sub(r"(?<!\r)\n", "\r\n", email)


def sanitizeTabs(email: str) -> str:            # sanitizers.ts:61-63 (first occurrence only, as String.replace does)
    return email.replace("=09", "\t", 1)


sanitizers = [revertGoogleMessageId, removeLabels, insert13Before10, sanitizeTabs]   # sanitizers.ts:65


class _Attempt:
    def __init__(self, result=None, comment=None, domain="", found=False):
        self.result, self.comment, self.domain, self.found = result, comment, domain, found


def _try_verify_dkim(raw: bytes, domain: str, skip_body_hash: bool, resolver) -> _Attempt:
    """tryVerifyDKIM (dkim/index.ts:99-158): the result for `domain` (default: the From domain)."""
    parsed, body = split_message(raw)
    if not domain:
        froms = [h for k, h in parsed if k == "from"]
        if len(froms) > 1:
            raise ValueError("Multiple From header in email and domain for verification not specified")
        m = re.search(rb"[\w.+-]+@([\w.-]+)", froms[0]) if froms else None
        domain = m.group(1).decode().lower() if m else ""
    last_reason = None
    found = False
    for key, line in parsed:
        if key != "dkim-signature":
            continue
        value = re.sub(r"\r?\n[ \t]*", " ", line.decode("latin-1").split(":", 1)[1])
        tags = parse_tag_list(value)
        if tags.get("d", "").lower() != domain:
            continue
        found = True
        algo = tags.get("a", "rsa-sha256").lower()
        if algo != "rsa-sha256":
            last_reason = f"unsupported algorithm {algo}"
            continue
        canon = tags.get("c", "simple/simple").lower()
        hc, bc = (canon.split("/") + ["simple"])[:2]
        canon_body = relaxed_body(body) if bc == "relaxed" else simple_body(body)
        if "l" in tags:
            canon_body = canon_body[: int(tags["l"])]
        body_hash = base64.b64encode(hashlib.sha256(canon_body).digest()).decode()
        bh_tag = re.sub(r"\s+", "", tags.get("bh", ""))
        if not skip_body_hash and body_hash != bh_tag:
            last_reason = "body hash did not verify"
            continue
        headers = signed_header_bytes(parsed, line, tags.get("h", ""), hc)
        try:
            signature = base64.b64decode(re.sub(r"\s+", "", tags.get("b", "")))
        except Exception:
            last_reason = "bad signature"
            continue
        selector = tags.get("s", "")
        try:
            records = resolver(f"{selector}._domainkey.{domain}", "TXT")
        except Exception:  # dkim/index.ts:105-131: no record from any source
            last_reason = "no key"
            continue
        tried = False
        for rec in records:     # every key the resolver knows for this selector is tried (DNS + archive keys, :120-129)
            ktags = parse_tag_list(rec)
            if "p" not in ktags or not ktags["p"]:
                continue
            try:
                pub = serialization.load_der_public_key(base64.b64decode(re.sub(r"\s+", "", ktags["p"])))
            except Exception:
                continue
            if not isinstance(pub, rsa.RSAPublicKey):
                continue
            tried = True
            try:
                pub.verify(signature, headers, padding.PKCS1v15(), hashes.SHA256())
            except InvalidSignature:
                last_reason = "bad signature"
                continue
            return _Attempt(DKIMVerificationResult(
                publicKey=pub.public_numbers().n, signature=int.from_bytes(signature, "big"), headers=headers,
                body=canon_body, bodyHash=bh_tag, signingDomain=domain, selector=selector, algo=algo,
                format=canon, modulusLength=pub.key_size), None, domain, True)
        if not tried:
            last_reason = last_reason or "no key"
    return _Attempt(None, last_reason, domain, found)


def verify_dkim_signature(email: bytes | str, domain: str = "", enable_sanitization: bool = True,
                          fallback_to_zk_email_dns_archive: bool = False, skip_body_hash: bool = False,
                          resolver=None) -> DKIMVerificationResult:
    """verifyDKIMSignature (dkim/index.ts:36-97).  `resolver(name, "TXT") -> list[str]` supplies DKIM key records."""
    raw = email.encode("latin-1") if isinstance(email, str) else bytes(email)
    resolver = resolver or _offline_resolver
    att = _try_verify_dkim(raw, domain, skip_body_hash, resolver)
    if not att.found:
        raise ValueError(f"DKIM signature not found for domain {att.domain}")
    applied = None
    if att.result is None and att.comment == "bad signature" and enable_sanitization:      # index.ts:49-66
        text = raw.decode("latin-1")
        for sanitize in sanitizers:
            retry = _try_verify_dkim(sanitize(text).encode("latin-1"), domain, skip_body_hash, resolver)
            if retry.result is not None:
                att, applied = retry, sanitize.__name__
                break
    if att.result is None:
        raise ValueError(f"DKIM signature verification failed for domain {att.domain}. Reason: {att.comment}")
    att.result.appliedSanitization = applied
    return att.result
