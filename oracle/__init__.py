"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

backend() returns the module that answers "what does the reference compute?":
  * oracle.refapi  -- the UNMODIFIED reference objects (oracle/_ref/liblgs_ref.so), preferred;
  * oracle.portapi -- the plain-C restatement (oracle/lgs_oracle.c), pinned against the former
                      and against tests/golden/ by tests/test_oracle_port.py.
Set LGS_ORACLE=port to force the restatement.
"""
import os


def backend():
    from . import portapi, refapi
    if os.environ.get("LGS_ORACLE", "") != "port" and refapi.available():
        return refapi
    return portapi


def backend_name() -> str:
    return backend().__name__.rsplit(".", 1)[-1]
