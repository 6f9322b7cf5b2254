"""TEST INFRASTRUCTURE ONLY: builds and loads the SIMT-emulator build of the kernel sources
(tests/emu/libsccg_b200_emu.so) so that kernel logic can be checked against the oracle on a
machine without a GPU.  Never used by the product path, bench.py or smoke()."""
import subprocess
from pathlib import Path

import sccg_b200

EMU_DIR = Path(__file__).resolve().parent / "emu"
EMU_LIB = EMU_DIR / "libsccg_b200_emu.so"


def emu_context() -> "sccg_b200.Context":
    subprocess.check_call(["sh", str(EMU_DIR / "build_emu.sh")], stdout=subprocess.DEVNULL)
    return sccg_b200.Context(0, lib_path=EMU_LIB)
