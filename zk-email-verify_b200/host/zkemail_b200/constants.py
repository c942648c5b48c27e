"""Mirror of /root/reference/packages/helpers/src/constants.ts:1-7."""
CIRCOM_FIELD_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MAX_HEADER_PADDED_BYTES = 1024  # default max size to be used in circuit
MAX_BODY_PADDED_BYTES = 1536    # default max size to be used in circuit
CIRCOM_BIGINT_N = 121
CIRCOM_BIGINT_K = 17
CIRCOM_LEVELS = 30
