"""Mirror of /root/reference/packages/helpers/src/input-generators.ts."""
from __future__ import annotations
from .binary_format import to_circom_bigint_bytes, uint8array_to_char_array
from .constants import MAX_BODY_PADDED_BYTES, MAX_HEADER_PADDED_BYTES
from .dkim import DKIMVerificationResult, verify_dkim_signature
from .sha_utils import generate_partial_sha, sha256_pad


def remove_soft_line_breaks(body: bytes):
    """removeSoftLineBreaks (input-generators.ts:130-159): returns (cleanContent padded to len(body), positionMap)."""
    result = bytearray()
    position_map = {}
    i = 0
    while i < len(body):
        if i + 2 < len(body) and body[i] == 61 and body[i + 1] == 13 and body[i + 2] == 10:
            i += 3
        else:
            position_map[len(result)] = i
            result.append(body[i])
            i += 1
    result.extend(b"\x00" * (len(body) - len(result)))
    return bytes(result), position_map


def _find_selector_in_clean_content(clean: bytes, selector: str, position_map: dict):
    clean_str = clean.decode("utf-8", errors="replace")
    idx = clean_str.find(selector)
    if idx == -1:
        raise ValueError(f'SHA precompute selector "{selector}" not found in cleaned body')
    if idx not in position_map:
        raise ValueError("Failed to map selector position to original body")
    return position_map[idx]


def _get_adjusted_selector(original_body: bytes, selector: str, clean: bytes, position_map: dict) -> str:
    """getAdjustedSelector (input-generators.ts:91-108)."""
    body_str = original_body.decode("utf-8", errors="replace")
    if selector in body_str:
        return selector
    original_index = _find_selector_in_clean_content(clean, selector, position_map)
    return body_str[original_index: original_index + len(selector) + 3]


def generate_email_verifier_inputs_from_dkim_result(dkim_result: DKIMVerificationResult, params: dict | None = None) -> dict:
    """generateEmailVerifierInputsFromDKIMResult (input-generators.ts:190-252).  `params` keys are the reference's
    InputGenerationArgs (input-generators.ts:20-30): ignoreBodyHashCheck, enableHeaderMasking, enableBodyMasking,
    shaPrecomputeSelector, maxHeadersLength, maxBodyLength, removeSoftLineBreaks, headerMask, bodyMask."""
    params = params or {}
    headers, body, body_hash = dkim_result.headers, dkim_result.body, dkim_result.bodyHash
    message_padded, message_padded_len = sha256_pad(headers, params.get("maxHeadersLength") or MAX_HEADER_PADDED_BYTES)
    circuit_inputs = {
        "emailHeader": uint8array_to_char_array(message_padded),
        "emailHeaderLength": str(message_padded_len),
        "pubkey": to_circom_bigint_bytes(dkim_result.publicKey),
        "signature": to_circom_bigint_bytes(dkim_result.signature),
    }
    if params.get("enableHeaderMasking"):
        circuit_inputs["headerMask"] = params.get("headerMask")
    if not params.get("ignoreBodyHashCheck"):
        if not body or not body_hash:
            raise ValueError("body and bodyHash are required when ignoreBodyHashCheck is false")
        body_hash_index = headers.decode("latin-1").find(body_hash)
        max_body_length = params.get("maxBodyLength") or MAX_BODY_PADDED_BYTES
        body_sha_length = ((len(body) + 63 + 65) // 64) * 64
        body_padded, body_padded_len = sha256_pad(body, max(max_body_length, body_sha_length))
        adjusted_selector = params.get("shaPrecomputeSelector")
        if adjusted_selector:
            clean, position_map = remove_soft_line_breaks(body_padded)
            adjusted_selector = _get_adjusted_selector(body, adjusted_selector, clean, position_map)
        precomputed_sha, body_remaining, body_remaining_length = generate_partial_sha(
            body_padded, body_padded_len, adjusted_selector, max_body_length)
        circuit_inputs["emailBodyLength"] = str(body_remaining_length)
        circuit_inputs["precomputedSHA"] = uint8array_to_char_array(precomputed_sha)
        circuit_inputs["bodyHashIndex"] = str(body_hash_index)
        circuit_inputs["emailBody"] = uint8array_to_char_array(body_remaining)
        if params.get("removeSoftLineBreaks"):
            clean, _ = remove_soft_line_breaks(body_remaining)
            circuit_inputs["decodedEmailBodyIn"] = uint8array_to_char_array(clean)
        if params.get("enableBodyMasking"):
            circuit_inputs["bodyMask"] = params.get("bodyMask")
    return circuit_inputs


def generate_email_verifier_inputs(raw_email: bytes | str, input_params: dict | None = None,
                                   dkim_verification_args: dict | None = None, resolver=None) -> dict:
    """generateEmailVerifierInputs (input-generators.ts:168-181)."""
    a = dkim_verification_args or {}
    dkim_result = verify_dkim_signature(raw_email, a.get("domain", ""), a.get("enableSanitization", True),
                                        a.get("fallbackToZKEmailDNSArchive", False), resolver=resolver)
    return generate_email_verifier_inputs_from_dkim_result(dkim_result, input_params)


# BASELINE.json's north_star uses the pre-rename name; keep it as an alias
generate_circuit_inputs = generate_email_verifier_inputs


TWITTER_SELECTOR = "email was meant for @"   # /root/reference/docs/zk-email-docs/UsageGuide/README.md:84


def generate_twitter_verifier_inputs_from_dkim_result(dkim_result: DKIMVerificationResult, address: int | str,
                                                      params: dict | None = None) -> dict:
    """Inputs of the Proof-of-Twitter circuit (BASELINE configs[3]; `TwitterVerifier` in capi_circuit.cpp): the
    EmailVerifier inputs with `shaPrecomputeSelector` = the Twitter selector (UsageGuide/README.md:84), plus
    `twitterUsernameIndex` = position of the user name in the remaining body and the public `address`.  The reference
    repo documents the selector and ships the resulting proof fixture; the generator script itself lives in the
    (un-vendored) proof-of-twitter app."""
    p = dict(params or {})
    p.setdefault("shaPrecomputeSelector", TWITTER_SELECTOR)
    inputs = generate_email_verifier_inputs_from_dkim_result(dkim_result, p)
    body = bytes(int(x) for x in inputs["emailBody"])
    at = body.find(TWITTER_SELECTOR.encode())
    if at < 0:
        raise ValueError(f'Sha precompute selector "{TWITTER_SELECTOR}" not found in the body')
    inputs["twitterUsernameIndex"] = str(at + len(TWITTER_SELECTOR))
    inputs["address"] = str(int(address, 0) if isinstance(address, str) else int(address))
    return inputs
