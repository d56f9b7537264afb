import faulthandler, sys
faulthandler.dump_traceback_later(45, exit=True)
import pytest
sys.exit(pytest.main(["tests/test_gpu_regrid.py", "-x", "-q", "-s"]))
