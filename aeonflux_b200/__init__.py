"""aeonflux_b200: B200-native batch engine for aeonflux's issuer hot path (Issuer::verify, CredentialIssuance::verify)."""
from .issuer import (KIND_PUBLIC_POINT, KIND_PUBLIC_SCALAR, KIND_SECRET_POINT, KIND_SECRET_SCALAR, IssuanceBatch, Issuer,  # noqa: F401
                     PresentationBatch, RequestBatch, compact_to_batchable)
