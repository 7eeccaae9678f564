"""TEST INFRASTRUCTURE (checker only): CPU restatement of the reference hot path (`oracle.oracle`)
and ctypes access to the unmodified reference build (`oracle.ref`).  Never imported by
lphash_b200/."""
