#!/usr/bin/env python3
"""Drop-in for reference ``utils/VStrains_PE_Inference.py``.

Install this file at the same path inside a VStrains checkout: the caller
(reference utils/VStrains_SPAdes.py:118-132) keeps running
``python <utils>/VStrains_PE_Inference.py -g G -o DIR -f FWD -r RVE -k K`` and keeps reading
``DIR/pe_info`` and ``DIR/st_info``; the work happens on the B200 through libvspe.so."""
import os
import sys

_ROOT = os.environ.get("VSPE_HOME") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from vstrains_b200.pe_inference import VspeError, main, reverse_seq, single_end_read_mapping  # noqa: E402,F401

if __name__ == "__main__":
    try:
        main()
    except VspeError as e:
        print(str(e), file=sys.stderr)
        sys.exit(1)
    sys.exit(0)
