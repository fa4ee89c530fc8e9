// ldpc_code.hpp -- host-side parity-check tables of one code (runtime replacement of the
// reference's compile-time enums, ArrayLDPCMacro.h:17-40, and of the members ReadH fills,
// ArrayLDPCMacro.h:170-172).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct ldpc_code {
    int n = 0, m = 0, edges = 0, dc_max = 0, dv_max = 0;
    std::vector<int> cdeg, vdeg;
    std::vector<int> clist;  // [m][dc_max], ascending per row, -1 padded
    std::vector<int> vlist;  // [n][dv_max], ascending per row, -1 padded
    // slot (position inside the check's row) of edge j of variable v: what the reference's
    // running addr_count resolves to (ArrayLDPC_Decoder.cpp:131-154)
    std::vector<int> vslot;  // [n][dv_max]
    // array-code provenance (0 when the code came from a file)
    int array_p = 0, array_rows = 0;
    double rate = 0.0;
};

// Generator equations in the reference's Format B (members of FP_Encoder, ArrayLDPCMacro.h:191-206)
struct ldpc_gen_device;  // device tables of the batched encoder (ldpc_encode.cu), created on first use
struct ldpc_gen {
    int n = 0, rows = 0;
    ldpc_gen_device *dev = nullptr;
    std::vector<int> flag;                 // ColumnFlag[n]
    std::vector<int> info_index, parity_index;
    std::vector<std::vector<int>> eq;      // G_mlist rows
};

namespace ldpc {

void set_error(const std::string &msg);
const char *last_error();

// all return 0 or a negative ldpc_status and fill `out`
int build_from_checks(int n, int m, const int *cdeg, const int *clist, int cstride, ldpc_code &out);
int load_file(const char *path, int format, ldpc_code &out);
int build_array(int p, int nrows, const int *row_mult, int ncols, const int *col_sel, int backward,
                ldpc_code &out);
int save_format_a(const ldpc_code &code, const char *path);
int load_generator(const char *path, ldpc_gen &out);
int encode(const ldpc_gen &g, const char *info, int info_len, uint8_t *codeword);
int derive_generator(const ldpc_code &code, const int *parity_cols, int nparity, ldpc_gen &out);
int save_generator(const ldpc_gen &g, const char *path);

}  // namespace ldpc
