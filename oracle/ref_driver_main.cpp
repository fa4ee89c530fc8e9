/*
 * ref_driver_main.cpp -- TEST INFRASTRUCTURE ONLY.
 * Single-TU build of the reference's own console program (Wrapper.cpp main + PerfTest.cpp
 * drivers) for one code variant, used to reproduce its printed transcripts
 * (wifi_results_4_4_2dB_30iter.txt; ArrayLDPC_PerfTest "3000 100 100").  See
 * ref_harness.cpp for how the variant header is substituted.  Output:
 * oracle/_ref/wrapper_<variant>.
 */
#include <fstream>
#include <iostream>
#include <math.h>
#include "ArrayLDPCMacro.h" /* variant header first; the guard silences the original */

#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#define REF_FILE(name) REF_STR(REF_DIR/name)

#include REF_FILE(rngs.cpp)
#include REF_FILE(rvgs.cpp)
#include REF_FILE(ArrayLDPC_Decoder.cpp)
#include REF_FILE(ArrayLDPC_Encoder.cpp)
#include REF_FILE(PerfTest.cpp)
#include REF_FILE(Wrapper.cpp)
