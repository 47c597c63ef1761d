"""TEST INFRASTRUCTURE ONLY — CPU oracles for the forced-alignment hot path.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs.  The product
package (whisperx_b200/) must never import anything from here.
"""
