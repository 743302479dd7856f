#!/usr/bin/env python
"""Drop-in for step 4 of the reference's full_pipeline_maMIMO_DNNEst.sh:44-48: replace

    $PY massiveMIMO_CSI_prediction_DNN.py --test -x ... --nn 1024 1024 -d ... --modeldir ... --useGPU 0 --useBN \
        --datasource matlab_maMimo --valSameTrain

by the same line with this file's name.  Flags, files written and exit codes follow the reference's --test branch;
the nets run on the B200 engine (see dl-channel-estimation-mamimo_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    import mamimo_b200 as mm
    sys.exit(mm.cli.main())
