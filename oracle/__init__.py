"""Parity oracle (TEST INFRASTRUCTURE ONLY -- never imported by fhmcanalysis_b200).

* ``oracle.fhmc_oracle``  -- NumPy restatement of the reference's hot path, each function citing
  the reference file:line it follows.
* ``oracle.ref``          -- loader for the compiled, unmodified-algorithm reference (oracle/_ref),
  built by ``oracle/build_ref.py`` where /root/reference exists.
"""
