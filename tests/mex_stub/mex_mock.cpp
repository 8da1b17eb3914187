// A minimal executable stand-in for MATLAB's libmx / libmex -- TEST INFRASTRUCTURE, not a MATLAB library.
//
// It implements exactly the part of the MEX C API that matlab/gpz_b200_mex.cpp uses (declared in mex.h next to this file):
// mxArray with double / logical / char / struct classes, column-major data, mexErrMsgIdAndTxt raising a C++ exception the way
// MATLAB unwinds a MEX call.  Together with the gateway source it is built into tests/mex_stub/libgpz_mex_harness.so, whose
// `mock_*` entry points let tests/test_mex_gateway.py build argument arrays, call mexFunction and read the results back, so
// that the gateway RUNS (create -> eval -> fit -> phi -> get_prior -> train -> predict -> destroy) instead of only compiling.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"

enum MockClass { MOCK_DOUBLE = 0, MOCK_LOGICAL = 1, MOCK_CHAR = 2, MOCK_STRUCT = 3 };

struct mxArray_tag {
    int cls = MOCK_DOUBLE;
    std::vector<size_t> dims;
    std::vector<double> real;
    std::vector<unsigned char> logical;       // mxLogical is bool: one byte
    std::string chars;
    std::map<std::string, mxArray*> fields;   // 1 x 1 struct
    size_t numel() const {
        size_t n = 1;
        for (size_t d : dims) n *= d;
        return dims.empty() ? 0 : n;
    }
};

struct MockError : std::runtime_error {
    std::string id;
    MockError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};

static void (*g_at_exit)(void) = nullptr;
static std::string g_printed;

extern "C" {

double* mxGetPr(const mxArray* a) { return a->cls == MOCK_DOUBLE && !a->real.empty() ? const_cast<double*>(a->real.data()) : nullptr; }
double mxGetScalar(const mxArray* a) {
    if (a == nullptr || a->numel() == 0) throw MockError("mock:scalar", "mxGetScalar of an empty array");
    if (a->cls == MOCK_LOGICAL) return a->logical[0] ? 1.0 : 0.0;
    if (a->cls == MOCK_CHAR) return static_cast<double>(a->chars[0]);
    return a->real[0];
}
size_t mxGetM(const mxArray* a) { return a->dims.empty() ? 0 : a->dims[0]; }
size_t mxGetN(const mxArray* a) {
    if (a->dims.size() < 2) return a->dims.empty() ? 0 : 1;
    size_t n = 1;
    for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
    return n;
}
size_t mxGetNumberOfElements(const mxArray* a) { return a->numel(); }
bool mxIsEmpty(const mxArray* a) { return a->numel() == 0; }
bool mxIsChar(const mxArray* a) { return a->cls == MOCK_CHAR; }
bool mxIsStruct(const mxArray* a) { return a->cls == MOCK_STRUCT; }
bool mxIsLogical(const mxArray* a) { return a->cls == MOCK_LOGICAL; }
mxLogical* mxGetLogicals(const mxArray* a) { return reinterpret_cast<mxLogical*>(const_cast<unsigned char*>(a->logical.data())); }
mxArray* mxGetField(const mxArray* a, size_t, const char* name) {
    auto it = a->fields.find(name);
    return it == a->fields.end() ? nullptr : it->second;
}
int mxGetString(const mxArray* a, char* buf, size_t len) {
    if (a == nullptr || a->cls != MOCK_CHAR || len == 0) return 1;
    const size_t n = a->chars.size() < len - 1 ? a->chars.size() : len - 1;
    std::memcpy(buf, a->chars.data(), n);
    buf[n] = 0;
    return a->chars.size() > len - 1;
}
mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity) {
    mxArray* a = new mxArray_tag;
    a->dims = {m, n};
    a->real.assign(m * n, 0.0);
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->real[0] = v;
    return a;
}
mxArray* mxCreateNumericArray(size_t nd, const mwSize* dims, mxClassID, mxComplexity) {
    mxArray* a = new mxArray_tag;
    a->dims.assign(dims, dims + nd);
    a->real.assign(a->numel(), 0.0);
    return a;
}
void mxDestroyArray(mxArray* a) {
    if (a == nullptr) return;
    for (auto& kv : a->fields) mxDestroyArray(kv.second);
    delete a;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw MockError(id, buf);
}
int mexPrintf(const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_printed += buf;
    return n;
}
int mexEvalString(const char*) { return 0; }
void mexLock(void) {}
int mexAtExit(void (*fn)(void)) {
    g_at_exit = fn;
    return 0;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

// ---------------------------------------------------------------------------------------------- harness (ctypes side)
mxArray* mock_double(const double* data, int nd, const int64_t* dims) {
    mxArray* a = new mxArray_tag;
    for (int i = 0; i < nd; ++i) a->dims.push_back(static_cast<size_t>(dims[i]));
    a->real.assign(data, data + a->numel());
    return a;
}
mxArray* mock_logical(const unsigned char* data, int64_t n) {
    mxArray* a = new mxArray_tag;
    a->cls = MOCK_LOGICAL;
    a->dims = {static_cast<size_t>(n), 1};
    a->logical.assign(data, data + n);
    return a;
}
mxArray* mock_string(const char* s) {
    mxArray* a = new mxArray_tag;
    a->cls = MOCK_CHAR;
    a->chars = s;
    a->dims = {1, a->chars.size()};
    return a;
}
mxArray* mock_struct(void) {
    mxArray* a = new mxArray_tag;
    a->cls = MOCK_STRUCT;
    a->dims = {1, 1};
    return a;
}
void mock_set_field(mxArray* s, const char* name, mxArray* v) {          // the struct takes ownership of v
    auto it = s->fields.find(name);
    if (it != s->fields.end()) mxDestroyArray(it->second);
    s->fields[name] = v;
}
mxArray* mock_empty(void) {
    mxArray* a = new mxArray_tag;
    a->dims = {0, 0};
    return a;
}
void mock_free(mxArray* a) { mxDestroyArray(a); }
int mock_ndim(const mxArray* a) { return static_cast<int>(a->dims.size()); }
int64_t mock_dim(const mxArray* a, int i) { return static_cast<int64_t>(a->dims[i]); }
const double* mock_data(const mxArray* a) { return a->real.data(); }
const char* mock_printed(void) { return g_printed.c_str(); }
void mock_clear_printed(void) { g_printed.clear(); }
void mock_run_at_exit(void) {
    if (g_at_exit) g_at_exit();
}

// calls the gateway; returns 0, or 1 with "<id>|<message>" in err when it raised (as MATLAB would turn into an error)
int mock_call(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs, char* err, int errlen) {
    try {
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray**>(prhs));
        return 0;
    } catch (const MockError& e) {
        snprintf(err, errlen, "%s|%s", e.id.c_str(), e.what());
        return 1;
    } catch (const std::exception& e) {
        snprintf(err, errlen, "c++|%s", e.what());
        return 1;
    }
}

}  // extern "C"
