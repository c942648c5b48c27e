"""B200-native host package mirroring the @zk-email/helpers surface for the EmailVerifier path
(/root/reference/packages/helpers/src/index.ts:1-4)."""
from .circuit import Circuit, FR_MODULUS  # noqa: F401
from .constants import *  # noqa: F401,F403
from .binary_format import (bigint_to_chunked_bytes, bytes_to_bigint, int64_to_bytes, int8_to_bytes,  # noqa: F401
                            to_circom_bigint_bytes, uint8array_to_char_array)
from .sha_utils import generate_partial_sha, partial_sha, sha256_pad, sha_hash  # noqa: F401
from .dkim import DKIMVerificationResult, verify_dkim_signature  # noqa: F401
from .input_generators import (generate_circuit_inputs, generate_email_verifier_inputs,  # noqa: F401
                               generate_twitter_verifier_inputs_from_dkim_result,
                               generate_email_verifier_inputs_from_dkim_result)
from .engine import AssertFailed, Context, Zkey, device_count, proof_to_json, verify, verify_batch  # noqa: F401
from .chunked_zkey import (generate_proof, verify_proof, register_circuit, register_zkey_files, generateProof, verifyProof,  # noqa: F401
                           InsecureKeyError)
from . import synthetic  # noqa: F401,E402
from . import iden3_binfile  # noqa: F401,E402
from . import verifier_args  # noqa: F401,E402
