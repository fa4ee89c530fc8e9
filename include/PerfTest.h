/* PerfTest.h -- the reference's driver entry points (PerfTest.h:4-11), implemented on the B200 engine in
 * fixedpointldpc_b200/csrc/facade/PerfTest.cpp.  Same names, arguments and console output; the frame loops run
 * as batched Monte-Carlo launches with the reference's own noise stream, so the printed numbers are the
 * reference's numbers. */
#ifndef PERF_TEST_H
#define PERF_TEST_H

void noMoreMemory();
int ArrayLDPC_Debug();
int ArrayLDPC_Debug_Wifi();
int ArrayLDPC_PerfTest(double db_start, double db_end, double db_step, char *Filename);
int ArrayLDPC_TimeTrial(double db, int MaxPckNum, char *Filename);
int DecodeTrial(double EbN0_dB, int MaxPacket);
int EncodeTrial(char *info, int MaxPacket);
int ArrayLDPC_Debug_Shorten(int short_len);

/* Extensions the reference stubs out: a real Eb/N0 sweep with CSV rows
 * "EbN0_dB,frames,frame_errors,bit_errors,FER,BER,avg_iters" appended to Filename and one line per point in
 * Filename_log.txt (histogram of the decoder's return values, GPUs, wall time).  short_len > 0: the first short_len
 * information positions (generator derived from H) are known zeros pinned to LLR 7*2^FRAC_WIDTH, like
 * ArrayLDPC_Debug_Shorten does for the array code (BASELINE config 3 on the cut79 variant). */
int ArrayLDPC_Sweep(double db_start, double db_end, double db_step, const char *Filename, int frame_errors, int short_len = 0);
/* State of the process-wide noise stream (rngs.cpp:45-49 keeps it in a file-static; default 123456789). */
void LDPC_PutSeed(long x);
long LDPC_GetSeed();
#endif
