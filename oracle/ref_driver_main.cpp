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
#include <stdlib.h>
#include <string.h>
#include "ArrayLDPCMacro.h" /* variant header first; the guard silences the original */

#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#define REF_FILE(name) REF_STR(REF_DIR/name)

#include REF_FILE(rngs.cpp)
#include REF_FILE(rvgs.cpp)
#include REF_FILE(ArrayLDPC_Decoder.cpp)
#include REF_FILE(ArrayLDPC_Encoder.cpp)
#include REF_FILE(PerfTest.cpp)
#define main ref_wrapper_main /* the reference's own main(): ArrayLDPC_PerfTest for argc == 5, else ArrayLDPC_Debug_Wifi */
#include REF_FILE(Wrapper.cpp)
#undef main

/* Wrapper.cpp leaves most drivers unreachable (Wrapper.cpp:34-99 is commented out or behind `return`); this entry
 * forwards to them by name and otherwise to the reference's own main.  Nothing here restates reference logic. */
int main(int argc, char *argv[])
{
    if (argc >= 3 && !strcmp(argv[1], "shorten")) return ArrayLDPC_Debug_Shorten(atoi(argv[2]));
    if (argc >= 2 && !strcmp(argv[1], "debug")) return ArrayLDPC_Debug();
    if (argc >= 4 && !strcmp(argv[1], "timetrial")) { char name[] = "timing.txt"; return ArrayLDPC_TimeTrial(atof(argv[2]), atoi(argv[3]), name); }
    return ref_wrapper_main(argc, argv);
}
