// bindings.cpp — python module cpd_b200._C.
//
// Deliberately torch-free: every entry point takes raw device/host addresses
// (tensor.data_ptr()), sizes and the CUDA stream handle as integers; shape, dtype,
// device and contiguity validation lives in cpd_b200/ops/native.py.  This keeps the
// native library independent of libtorch's C++ ABI (it links only the CUDA runtime)
// and makes it usable from any framework that can hand out pointers.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <cstring>
#include <stdexcept>

#include "api.h"
#include "comm/comm.h"
#include "common/fpformat.h"

namespace py = pybind11;
using uptr = uintptr_t;

namespace {

template <class T> T* P(uptr p) { return reinterpret_cast<T*>(p); }
cudaStream_t S(uptr s) { return reinterpret_cast<cudaStream_t>(s); }

void check_fmt(int e, int m) {
  if (e < 1 || e > 8) throw std::invalid_argument("exp must be in [1, 8]");
  if (m < 0 || m > 23) throw std::invalid_argument("man must be in [0, 23]");
}
// every kernel launch of this library passes through check_launch(): the counter is what bench.py reports as
// `gpu_launches` (counted, not estimated; launches recorded into a CUDA graph count once per capture)
std::atomic<long long> g_launches{0};
void check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) throw std::runtime_error(std::string("cpd_b200: ") + what + " launch failed: " + cudaGetErrorString(err));
}

struct PyCommBuffers { cpd::CommBuffers cb; };

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.doc() = "cpd_b200 native ops (sm_100a CUDA kernels, C++ CPU reference paths, comm runtime)";

  // ---------------------------------------------------------------- elementwise --
  m.def("quantize_cpu", [](uptr in, uptr out, int64_t n, int e, int mb, int rounding, int shift) {
    check_fmt(e, mb);
    py::gil_scoped_release nogil;
    if (shift == 0) cpd::quantize_cpu(P<const float>(in), P<float>(out), n, e, mb, rounding);
    else cpd::quantize_scaled_cpu(P<const float>(in), P<float>(out), n, e, mb, rounding, shift);
  });
  m.def("quantize_cuda", [](uptr in, uptr out, int64_t n, int e, int mb, int rounding, int shift, uptr stream) {
    check_fmt(e, mb);
    cpd::quantize_cuda(P<const float>(in), P<float>(out), n, e, mb, rounding, shift, S(stream));
    check_launch("quantize");
  });
  m.def("amax_cuda", [](uptr in, int64_t n, uptr amax_bits, uptr stream) {
    cpd::amax_cuda(P<const float>(in), n, P<unsigned>(amax_bits), S(stream));
    check_launch("amax");
  });

  // ----------------------------------------------------------------------- GEMM --
  m.def("gemm_exact_cpu", [](uptr A, int64_t sa0, int64_t sa1, uptr B, int64_t sb0, int64_t sb1, uptr C, int64_t ldc, int M,
                             int N, int K, int e, int mb, int rounding, bool kahan) {
    check_fmt(e, mb);
    py::gil_scoped_release nogil;
    cpd::gemm_exact_cpu(P<const float>(A), sa0, sa1, P<const float>(B), sb0, sb1, P<float>(C), ldc, M, N, K, e, mb, rounding, kahan);
  });
  m.def("gemm_tile_cpu", [](uptr A, int64_t sa0, int64_t sa1, uptr B, int64_t sb0, int64_t sb1, uptr C, int64_t ldc, int M,
                            int N, int K, int e, int mb, int rounding, bool kahan, int k_slice, int op_e, int op_m, int splits) {
    check_fmt(e, mb);
    py::gil_scoped_release nogil;
    cpd::gemm_tile_cpu(P<const float>(A), sa0, sa1, P<const float>(B), sb0, sb1, P<float>(C), ldc, M, N, K, e, mb, rounding, kahan,
                       k_slice, op_e, op_m, splits);
  });
  m.def("gemm_exact_cuda", [](uptr A, int64_t sa0, int64_t sa1, uptr B, int64_t sb0, int64_t sb1, uptr C, int64_t ldc, int M,
                              int N, int K, int e, int mb, int rounding, bool kahan, uptr stream) {
    check_fmt(e, mb);
    cpd::gemm_exact_cuda(P<const float>(A), sa0, sa1, P<const float>(B), sb0, sb1, P<float>(C), ldc, M, N, K, e, mb, rounding,
                         kahan, S(stream));
    check_launch("gemm_exact");
  });
  // A [M,K] row-major, Bt [N,K] row-major (both K-major); operand_mode 0 bf16, 1 tf32(fp32 in), 2 e4m3, 3 e5m2
  m.def("gemm_tc_cuda", [](uptr A, uptr Bt, uptr C, int M, int N, int K, int e, int mb, bool kahan, int k_slice,
                           int operand_mode, uptr bias, int variant, uptr stream, int64_t lda, int64_t ldb, int splits,
                           uptr workspace) {
    check_fmt(e, mb);
    int rc = cpd::gemm_tc_cuda(P<const void>(A), P<const void>(Bt), P<float>(C), M, N, K, e, mb, kahan, k_slice, operand_mode,
                               P<const float>(bias), variant, S(stream), lda, ldb, splits, P<float>(workspace));
    if (splits > 1 && rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);      // + the split-K fold kernel
    check_launch("gemm_tc");
    return rc;
  }, py::arg("A"), py::arg("Bt"), py::arg("C"), py::arg("M"), py::arg("N"), py::arg("K"), py::arg("e"), py::arg("m"),
     py::arg("kahan"), py::arg("k_slice"), py::arg("operand_mode"), py::arg("bias"), py::arg("variant"), py::arg("stream"),
     py::arg("lda") = 0, py::arg("ldb") = 0, py::arg("splits") = 1, py::arg("workspace") = (uptr)0);

  // implicit-GEMM convolution: X is the NHWC activation tensor in the operand dtype (read through an im2col-mode TMA
  // descriptor), Bt [N, KH*KW*C] the filter matrix in (kh, kw, c) order; geom = (B, H, W, C, KH, KW, OH, OW, stride, pad, dil)
  m.def("gemm_conv_cuda", [](uptr X, uptr Bt, uptr C, int N, std::vector<int> geom, int e, int mb, bool kahan, int k_slice,
                             int operand_mode, uptr bias, int variant, uptr stream, int64_t ldb, int splits, uptr workspace) {
    check_fmt(e, mb);
    if (geom.size() != 11) throw std::invalid_argument("gemm_conv: geom = (B, H, W, C, KH, KW, OH, OW, stride, pad, dil)");
    cpd::ConvGeom g{geom[0], geom[1], geom[2], geom[3], geom[4], geom[5], geom[6], geom[7], geom[8], geom[9], geom[10]};
    const long long M = (long long)g.B * g.OH * g.OW, K = (long long)g.KH * g.KW * g.C;
    if (M > INT32_MAX || K > INT32_MAX) throw std::invalid_argument("gemm_conv: problem too large");
    int rc = cpd::gemm_tc_cuda(P<const void>(X), P<const void>(Bt), P<float>(C), (int)M, N, (int)K, e, mb, kahan, k_slice, operand_mode,
                               P<const float>(bias), variant, S(stream), 0, ldb, splits, P<float>(workspace), &g);
    if (splits > 1 && rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);      // + the split-K fold kernel
    check_launch("gemm_conv");
    return rc;
  }, py::arg("X"), py::arg("Bt"), py::arg("C"), py::arg("N"), py::arg("geom"), py::arg("e"), py::arg("m"), py::arg("kahan"),
     py::arg("k_slice"), py::arg("operand_mode"), py::arg("bias"), py::arg("variant"), py::arg("stream"), py::arg("ldb") = 0,
     py::arg("splits") = 1, py::arg("workspace") = (uptr)0);

  m.def("gemm_wgrad_cuda", [](uptr dY, uptr X, uptr dW, int O, std::vector<int> geom, int e, int mb, bool kahan, int k_slice,
                              uptr stream, int splits, uptr workspace, int lbo, int sbo) {
    check_fmt(e, mb);
    if (geom.size() != 11) throw std::invalid_argument("gemm_wgrad: geom = (B, H, W, C, KH, KW, OH, OW, stride, pad, dil)");
    cpd::ConvGeom g{geom[0], geom[1], geom[2], geom[3], geom[4], geom[5], geom[6], geom[7], geom[8], geom[9], geom[10]};
    int rc = cpd::gemm_wgrad_cuda(P<const void>(dY), P<const void>(X), P<float>(dW), g, O, e, mb, kahan, k_slice, S(stream), splits,
                                  P<float>(workspace), lbo, sbo);
    if (splits > 1 && rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);      // + the split-K fold kernel
    check_launch("gemm_wgrad");
    return rc;
  }, py::arg("dY"), py::arg("X"), py::arg("dW"), py::arg("O"), py::arg("geom"), py::arg("e"), py::arg("m"), py::arg("kahan"),
     py::arg("k_slice"), py::arg("stream"), py::arg("splits") = 1, py::arg("workspace") = (uptr)0, py::arg("lbo") = 0, py::arg("sbo") = 0);

  // operand staging: fp32 -> tensor-core operand (gather / im2col + cast), see cuda/operand_stage_sm100.cu.
  // Descriptors are built once (opaque bytes) and one launch stages one or BOTH operands of a GEMM.
  m.def("stage_gather_desc", [](uptr src, int64_t base, int rows_fast, int64_t s_row, int64_t s1, int64_t s2, int64_t s3, int n2, int n3) {
    cpd::StageDesc d{};
    d.src = P<const float>(src); d.base = base; d.mode = 0; d.rows_fast = rows_fast;
    d.s_row = s_row; d.s1 = s1; d.s2 = s2; d.s3 = s3; d.n2 = n2 < 1 ? 1 : n2; d.n3 = n3 < 1 ? 1 : n3;
    return py::bytes(reinterpret_cast<const char*>(&d), sizeof(d));
  });
  m.def("stage_im2col_desc", [](uptr src, int64_t base, int rows_fast, int taps_are_rows, int C, int H, int W, int OH, int OW, int KH, int KW,
                                int stride, int pad, int dil, int in_dil, int64_t sb, int64_t sh, int64_t sw, int64_t sc) {
    cpd::StageDesc d{};
    d.src = P<const float>(src); d.base = base; d.mode = 1; d.rows_fast = rows_fast; d.taps_are_rows = taps_are_rows;
    d.C = C; d.H = H; d.W = W; d.OH = OH; d.OW = OW; d.KH = KH; d.KW = KW; d.stride = stride; d.pad = pad; d.dil = dil;
    d.in_dil = in_dil < 1 ? 1 : in_dil; d.sb = sb; d.sh = sh; d.sw = sw; d.sc = sc; d.n2 = d.n3 = 1;
    return py::bytes(reinterpret_cast<const char*>(&d), sizeof(d));
  });
  m.def("stage_launch", [](std::vector<std::string> descs, std::vector<uptr> dsts, std::vector<int64_t> R, std::vector<int64_t> K,
                           std::vector<int64_t> Kp, int operand_mode, uptr stream) {
    const size_t n = descs.size();
    if (n < 1 || n > 2 || dsts.size() != n || R.size() != n || K.size() != n || Kp.size() != n)
      throw std::invalid_argument("stage_launch: one or two operands");
    cpd::StageDesc d[2];
    void* out[2];
    long long r[2], k[2], kp[2];
    for (size_t i = 0; i < n; ++i) {
      if (descs[i].size() != sizeof(cpd::StageDesc)) throw std::invalid_argument("stage_launch: bad descriptor");
      std::memcpy(&d[i], descs[i].data(), sizeof(cpd::StageDesc));
      out[i] = P<void>(dsts[i]); r[i] = R[i]; k[i] = K[i]; kp[i] = Kp[i];
    }
    int rc = cpd::operand_stage2_cuda(d, out, r, k, kp, (int)n, operand_mode, S(stream));
    check_launch("operand_stage");
    return rc;
  });
  // fused BatchNorm + residual + ReLU (training), channels_last fp32
  m.def("bn_workspace_floats", [](int64_t M, int C) { return (int64_t)cpd::bn_workspace_floats(M, C); });
  m.def("bn_act_forward_cuda", [](uptr x, uptr res, uptr y, int64_t M, int C, uptr gamma, uptr beta, uptr rmean, uptr rvar, uptr nbt,
                                  double momentum, double eps, bool relu, uptr stats, uptr ws, uptr relu_mask, uptr stream) {
    int rc = cpd::bn_act_forward_cuda(P<const float>(x), P<const float>(res), P<float>(y), M, C, P<const float>(gamma), P<const float>(beta),
                                      P<float>(rmean), P<float>(rvar), P<long long>(nbt), (float)momentum, (float)eps, relu, P<float>(stats),
                                      P<float>(ws), P<unsigned>(relu_mask), S(stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);            // two kernels per call
    check_launch("bn_act_forward");
    return rc;
  });
  m.def("bn_act_backward_cuda", [](uptr dy, uptr y, uptr relu_mask, uptr x, uptr dx, uptr dres, int64_t M, int C, uptr gamma, uptr stats,
                                   bool relu, uptr grads, uptr ws, uptr stream) {
    int rc = cpd::bn_act_backward_cuda(P<const float>(dy), P<const float>(y), P<const unsigned>(relu_mask), P<const float>(x), P<float>(dx), P<float>(dres), M, C,
                                       P<const float>(gamma), P<const float>(stats), relu, P<float>(grads), P<float>(ws), S(stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    check_launch("bn_act_backward");
    return rc;
  });
  m.def("colsum_quantize_cuda", [](uptr x, int64_t rows, int cols, int64_t ld, uptr out, int e, int mb, int rounding, uptr stream) {
    check_fmt(e, mb);
    cpd::colsum_quantize_cuda(P<const float>(x), rows, cols, ld, P<float>(out), e, mb, rounding, S(stream));
    check_launch("colsum_quantize");
  });

  m.def("mx_quantize_cuda", [](uptr in, int64_t ld_in, uptr codes, uptr sf, int R, int K, int Kp, bool e5m2, uptr stream) {
    cpd::mx_quantize_cuda(P<const float>(in), ld_in, P<uint8_t>(codes), P<uint8_t>(sf), R, K, Kp, e5m2, S(stream));
    check_launch("mx_quantize");
  });
  m.def("gemm_tc_mx_cuda", [](uptr ca, uptr sa, uptr cb, uptr sb, uptr C, int M, int N, int Kp, int e, int mb, bool kahan, int k_slice,
                              bool e5m2, uptr bias, uptr stream) {
    check_fmt(e, mb);
    int rc = cpd::gemm_tc_mx_cuda(P<const uint8_t>(ca), P<const uint8_t>(sa), P<const uint8_t>(cb), P<const uint8_t>(sb), P<float>(C), M,
                                  N, Kp, e, mb, kahan, k_slice, e5m2, P<const float>(bias), S(stream));
    check_launch("gemm_tc_mx");
    return rc;
  });

  // --------------------------------------------------------- multi-tensor arenas --
  m.def("aps_fold_cpu", [](std::vector<uptr> srcs, uptr out, uptr seg_off, int T, int e, int mb, int rounding, bool use_aps, bool kahan,
                           double world_mult, uptr shifts_out) {
    check_fmt(e, mb);
    std::vector<const float*> ptrs;
    for (auto s : srcs) ptrs.push_back(P<const float>(s));
    py::gil_scoped_release nogil;
    cpd::aps_fold_cpu(ptrs.data(), (int)ptrs.size(), P<float>(out), P<const int64_t>(seg_off), T, e, mb, rounding, use_aps,
                      kahan, (float)world_mult, P<int>(shifts_out));
  });
  m.def("aps_fold_cuda", [](std::vector<uptr> srcs, uptr out, uptr chunks, int n_chunks, int T, int T_cap, int e, int mb, int rounding,
                            bool use_aps, bool kahan, double world_mult, uptr workspace, uptr shifts_out, int max_ctas, uptr stream) {
    check_fmt(e, mb);
    if (srcs.empty() || srcs.size() > (size_t)cpd::kMaxLocalSources) throw std::invalid_argument("aps_fold: 1..64 sources");
    std::vector<const float*> ptrs;
    for (auto s : srcs) ptrs.push_back(P<const float>(s));
    cpd::aps_fold_local_cuda(ptrs.data(), (int)ptrs.size(), P<float>(out), P<const cpd::Chunk>(chunks), n_chunks, T, T_cap, e, mb, rounding, use_aps,
                             kahan, (float)world_mult, P<unsigned>(workspace), P<int>(shifts_out), max_ctas, S(stream));
    check_launch("aps_fold");
  });
  m.def("lars_step_cpu", [](uptr master, uptr model, uptr grad, uptr mom, uptr seg_off, int T, double lr, double momentum,
                            double wd, double coeff, double eps, bool use_lars, bool nesterov, uptr wd_mult) {
    py::gil_scoped_release nogil;
    cpd::lars_step_cpu(P<float>(master), P<float>(model), P<const float>(grad), P<float>(mom), P<const int64_t>(seg_off), T,
                       (float)lr, (float)momentum, (float)wd, (float)coeff, (float)eps, use_lars, nesterov, P<const float>(wd_mult));
  });
  m.def("lars_step_cuda", [](uptr master, uptr model, uptr grad, uptr mom, uptr chunks, uptr tchunk_off, int n_chunks, int T,
                             uptr lr_dev, double lr, double momentum, double wd, double coeff, double eps, bool use_lars,
                             bool nesterov, uptr wd_mult, uptr partials, uptr barrier_ws, int max_ctas, uptr stream) {
    cpd::lars_step_cuda(P<float>(master), P<float>(model), P<const float>(grad), P<float>(mom), P<const cpd::Chunk>(chunks),
                        P<const int>(tchunk_off), n_chunks, T, P<const float>(lr_dev), (float)lr, (float)momentum, (float)wd,
                        (float)coeff, (float)eps, use_lars, nesterov, P<const float>(wd_mult), P<float>(partials),
                        P<unsigned>(barrier_ws), max_ctas, S(stream));
    check_launch("lars_step");
  });

  // ------------------------------------------------------------------ wire codec --
  m.def("wire_roundtrip_cpu", [](uptr in, uptr out, int64_t n, int e, int mb) {
    const float* p = P<const float>(in);
    float* o = P<float>(out);
    for (int64_t i = 0; i < n; ++i) o[i] = cpd::wire_unpack(cpd::wire_pack(p[i], e, cpd::wire_man(mb)), e, cpd::wire_man(mb));
  });
  m.def("wire_bits", [](int e, int mb) { return cpd::wire_bits(e, mb); });
  m.def("chunk_struct_size", []() { return (int)sizeof(cpd::Chunk); });
  m.def("sm_count", [](int dev) { return cpd::device_sm_count(dev); });
  m.def("launch_count", []() { return g_launches.load(std::memory_order_relaxed); },
        "number of CUDA kernel-launching entry points of this library called so far in this process");

  // ------------------------------------------------------------------------ comm --
  m.def("discover_topology_json", &cpd::comm::discover_topology_json);

  py::class_<cpd::comm::SymmRegion>(m, "SymmRegion")
      .def(py::init<size_t, int, int, int, const std::string&, bool>())
      .def("open", &cpd::comm::SymmRegion::open, py::call_guard<py::gil_scoped_release>())
      .def("exchange", &cpd::comm::SymmRegion::exchange, py::call_guard<py::gil_scoped_release>())
      .def("bind_multicast", &cpd::comm::SymmRegion::bind_multicast, py::call_guard<py::gil_scoped_release>())
      .def("close", &cpd::comm::SymmRegion::close_region, py::call_guard<py::gil_scoped_release>())
      .def("ptr", [](const cpd::comm::SymmRegion& r, int i) { return reinterpret_cast<uptr>(r.ptr(i)); })
      .def("mc_ptr", [](const cpd::comm::SymmRegion& r) { return reinterpret_cast<uptr>(r.mc_ptr()); })
      .def("bytes", &cpd::comm::SymmRegion::bytes)
      .def("log", &cpd::comm::SymmRegion::log);

  m.def("ipc_alloc", [](size_t bytes) { return reinterpret_cast<uptr>(cpd::comm::IpcRegion::alloc(bytes)); });
  m.def("ipc_export", [](uptr p) { return py::bytes(cpd::comm::IpcRegion::export_handle(P<void>(p))); });
  m.def("ipc_open", [](const std::string& h) { return reinterpret_cast<uptr>(cpd::comm::IpcRegion::open_handle(h)); });
  m.def("ipc_close", [](uptr p) { cpd::comm::IpcRegion::close_handle(P<void>(p)); });
  m.def("ipc_free", [](uptr p) { cpd::comm::IpcRegion::free_local(P<void>(p)); });

  py::class_<PyCommBuffers>(m, "CommBuffers")
      .def(py::init([](int rank, int world, std::vector<uptr> data, std::vector<uptr> gather, std::vector<uptr> mailbox,
                       std::vector<uptr> flags, uptr mc_gather, size_t cap_bytes, int max_tensors, uptr mc_data) {
        if (world < 1 || world > 8) throw std::invalid_argument("fused all-reduce supports 1..8 ranks per node");
        if ((int)data.size() != world || (int)gather.size() != world || (int)mailbox.size() != world || (int)flags.size() != world)
          throw std::invalid_argument("CommBuffers: need one pointer per rank");
        PyCommBuffers b;
        b.cb.rank = rank; b.cb.world = world;
        for (int r = 0; r < 16; ++r) { b.cb.data[r] = nullptr; b.cb.gather[r] = nullptr; b.cb.mailbox[r] = nullptr; b.cb.flags[r] = nullptr; }
        for (int r = 0; r < world; ++r) {
          b.cb.data[r] = P<unsigned char>(data[r]);
          b.cb.gather[r] = P<unsigned char>(gather[r]);
          b.cb.mailbox[r] = P<float>(mailbox[r]);
          b.cb.flags[r] = P<unsigned>(flags[r]);
        }
        b.cb.mc_gather = P<unsigned char>(mc_gather);
        b.cb.mc_data = P<unsigned char>(mc_data);
        b.cb.cap_bytes = cap_bytes;
        b.cb.max_tensors = max_tensors;
        return b;
      }), py::arg("rank"), py::arg("world"), py::arg("data"), py::arg("gather"), py::arg("mailbox"), py::arg("flags"),
         py::arg("mc_gather"), py::arg("cap_bytes"), py::arg("max_tensors"), py::arg("mc_data") = (uptr)0)
      .def_property_readonly("cap_bytes", [](const PyCommBuffers& b) { return b.cb.cap_bytes; })
      .def_property_readonly("max_tensors", [](const PyCommBuffers& b) { return b.cb.max_tensors; })
      .def_property_readonly("has_multicast", [](const PyCommBuffers& b) { return b.cb.mc_gather != nullptr; });

  m.def("allreduce_fused_cuda", [](const PyCommBuffers& b, uptr grad, int64_t numel, uptr chunks, int n_chunks, int T, int e, int mb,
                                   int rounding, bool use_aps, bool kahan, int algo, bool prequant, int wire, uptr local_ws, uptr shifts_out,
                                   int max_ctas, uint64_t timeout_cycles, uptr stream) {
    check_fmt(e, mb);
    if (T > 1024) throw std::invalid_argument("at most 1024 tensors per bucket");
    if (T > b.cb.max_tensors) throw std::invalid_argument("bucket has more tensors than the mailbox holds");
    if (wire != 8 && wire != 16 && wire != 32) throw std::invalid_argument("wire must be 8, 16 or 32");
    const int base_algo = algo & 0xFF;           // the upper bits carry opt-in protocol switches
    const bool fp32_wire = base_algo == cpd::ALGO_ONESHOT_RAW || base_algo == cpd::ALGO_NVLS_SUM;
    if ((size_t)numel * (size_t)(fp32_wire ? 4 : wire / 8) > b.cb.cap_bytes)
      throw std::invalid_argument("bucket larger than the symmetric wire buffer");
    if (n_chunks < 0 || (int64_t)n_chunks * 4096 < numel)
      throw std::invalid_argument("chunk table does not cover the bucket");
    cpd::allreduce_fused_cuda(b.cb, P<float>(grad), P<const cpd::Chunk>(chunks), n_chunks, T, e, mb, rounding, use_aps, kahan, algo, prequant,
                              wire, P<unsigned>(local_ws), P<int>(shifts_out), max_ctas, timeout_cycles, S(stream));
    check_launch("allreduce_fused");
  });
}
