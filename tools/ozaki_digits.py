"""Error study of the INT8 digit scheme (see oracle/int8_gram_oracle.py): python tools/ozaki_digits.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.int8_gram_oracle import fuzzed_and_plain

if __name__ == "__main__":
    fuzzed_and_plain()
