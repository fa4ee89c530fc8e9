// Wrapper.cpp (facade) -- console entry point with the reference's behaviour (Wrapper.cpp:17-101) and real
// arguments for the drivers the reference leaves unreachable.
//
//   wrapper <db_start> <db_end> <db_step> <csv>   ArrayLDPC_PerfTest           (argc == 5, Wrapper.cpp:23-28)
//   wrapper                                       ArrayLDPC_Debug_Wifi, Eb/N0 read from stdin (Wrapper.cpp:29-33)
//   wrapper debug | shorten <len> | decodetrial <dB> <frames> | encodetrial <frames> | timetrial <dB> <frames>
//   wrapper sweep <db_start> <db_end> <db_step> <csv> [frame_errors] [short_len]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>

#include "ArrayLDPCMacro.h"
#include "ArrayLDPC.h"
#include "PerfTest.h"

int main(int argc, char *argv[])
{
    try {
        if (argc >= 2 && !strcmp(argv[1], "debug")) return ArrayLDPC_Debug();
        if (argc >= 3 && !strcmp(argv[1], "shorten")) return ArrayLDPC_Debug_Shorten(atoi(argv[2]));
        if (argc >= 4 && !strcmp(argv[1], "decodetrial")) return DecodeTrial(atof(argv[2]), atoi(argv[3]));
        if (argc >= 4 && !strcmp(argv[1], "timetrial")) {
            char name[] = "timing.txt";
            return ArrayLDPC_TimeTrial(atof(argv[2]), atoi(argv[3]), name);
        }
        if (argc >= 3 && !strcmp(argv[1], "encodetrial")) {
            char info[248] = "OMG how long should this string be to make it 248";
            return EncodeTrial(info, atoi(argv[2]));
        }
        if (argc >= 6 && !strcmp(argv[1], "sweep"))
            return ArrayLDPC_Sweep(atof(argv[2]), atof(argv[3]), atof(argv[4]), argv[5], argc >= 7 ? atoi(argv[6]) : 100,
                                   argc >= 8 ? atoi(argv[7]) : 0);
        if (argc == 5) {
            // the reference parses nothing and always runs (2, 2, 1, "test.csv"); the arguments are honoured here
            return ArrayLDPC_PerfTest(atof(argv[1]), atof(argv[2]), atof(argv[3]), argv[4]);
        }
        return ArrayLDPC_Debug_Wifi();
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
}
