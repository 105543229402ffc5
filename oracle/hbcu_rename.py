"""Prints -Dhbcu_x=oracle_hbcu_x for every function include/hbcu.h declares (used by oracle/Makefile, libhostlogic.so)."""
import re
from pathlib import Path

text = (Path(__file__).resolve().parent.parent / "include" / "hbcu.h").read_text()
text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
print(" ".join(f"-D{f}=oracle_{f}" for f in sorted(set(re.findall(r"\b(hbcu_[a-z0-9_]+)\s*\(", text)))))
