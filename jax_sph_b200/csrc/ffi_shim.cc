// ffi_shim.cc -- XLA FFI (jax.ffi) handlers over the C ABI of libsphb200.so.
//
// NOT built by __graft_entry__.build(): the image has no jaxlib, hence no
// xla/ffi/api/ffi.h.  Where jax >= 0.4.38 is installed:
//
//   g++ -std=c++17 -O2 -fPIC -shared ffi_shim.cc -I$(python -c "import jax.ffi; \
//       print(jax.ffi.include_dir())") -I../../include -I/usr/local/cuda/include \
//       -L.. -lsphb200 -Wl,-rpath,'$ORIGIN' -o ../libsphb200_ffi.so
//
// and jax_sph_b200/jax_ffi.py registers the handlers.  The shim holds no logic:
// it unpacks XLA buffers into sphb200_state and forwards to sphb200_advance /
// sphb200_forward / sphb200_neighbor_list on XLA's stream, with the workspace
// taken from XLA's scratch allocator (sized by sphb200_workspace_bytes, a
// static function of the config and N, so shapes stay static under jit).
//
// Replaces: the body of `advance` (jax_sph/integrator.py:22-56) inside
// `jit(si_euler(...))` (jax_sph/simulate.py:91-93).
#include <cstdint>
#include <cstring>

#include "sphb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using I32 = ffi::Buffer<ffi::S32>;
using RF32 = ffi::Result<ffi::Buffer<ffi::F32>>;
using RI32 = ffi::Result<ffi::Buffer<ffi::S32>>;
using RU32 = ffi::Result<ffi::Buffer<ffi::U32>>;

ffi::Error Status(int rc) {
  if (rc == SPHB200_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, sphb200_strerror(rc));
}

// The config travels as one string attribute: the bytes of sphb200_config as lower-case hex
// (a str attribute decodes as std::string_view; an ndarray one would decode as a Span).
ffi::Error LoadConfig(std::string_view blob, sphb200_config* cfg) {
  if (blob.size() != 2 * sizeof(sphb200_config))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "sphb200_config size mismatch");
  auto nib = [](char ch) -> int {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    return -1;
  };
  unsigned char* dst = reinterpret_cast<unsigned char*>(cfg);
  for (size_t i = 0; i < sizeof(*cfg); ++i) {
    const int hi = nib(blob[2 * i]), lo = nib(blob[2 * i + 1]);
    if (hi < 0 || lo < 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "sphb200_config: not hex");
    dst[i] = static_cast<unsigned char>(hi * 16 + lo);
  }
  return ffi::Error::Success();
}

// advance(dt, state) -> state, err.  16 inputs / 16 outputs in the key order of
// solver.py:930-947; XLA may alias them (input_output_aliases in jax_ffi.py).
ffi::Error AdvanceImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, std::string_view cfg_blob,
                       double dt, F32 r, I32 tag, F32 u, F32 v, F32 dudt, F32 dvdt, F32 drhodt,
                       F32 rho, F32 p, F32 mass, F32 eta, F32 dTdt, F32 T, F32 kappa, F32 Cp,
                       F32 nw, RF32 r_o, RI32 tag_o, RF32 u_o, RF32 v_o, RF32 dudt_o, RF32 dvdt_o,
                       RF32 drhodt_o, RF32 rho_o, RF32 p_o, RF32 mass_o, RF32 eta_o, RF32 dTdt_o,
                       RF32 T_o, RF32 kappa_o, RF32 Cp_o, RF32 nw_o, RU32 err) {
  sphb200_config cfg;
  if (auto e = LoadConfig(cfg_blob, &cfg); e.failure()) return e;
  const int64_t n = r.dimensions()[0];
  size_t bytes = 0;
  if (int rc = sphb200_workspace_bytes(&cfg, n, &bytes)) return Status(rc);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "sphb200 workspace");
  sphb200_state in{}, out{};
  in.r = r.typed_data(); in.u = u.typed_data(); in.v = v.typed_data();
  in.dudt = dudt.typed_data(); in.dvdt = dvdt.typed_data(); in.nw = nw.typed_data();
  in.rho = rho.typed_data(); in.p = p.typed_data(); in.drhodt = drhodt.typed_data();
  in.mass = mass.typed_data(); in.eta = eta.typed_data(); in.T = T.typed_data();
  in.dTdt = dTdt.typed_data(); in.kappa = kappa.typed_data(); in.Cp = Cp.typed_data();
  in.tag = tag.typed_data();
  out.r = r_o->typed_data(); out.u = u_o->typed_data(); out.v = v_o->typed_data();
  out.dudt = dudt_o->typed_data(); out.dvdt = dvdt_o->typed_data(); out.nw = nw_o->typed_data();
  out.rho = rho_o->typed_data(); out.p = p_o->typed_data(); out.drhodt = drhodt_o->typed_data();
  out.mass = mass_o->typed_data(); out.eta = eta_o->typed_data(); out.T = T_o->typed_data();
  out.dTdt = dTdt_o->typed_data(); out.kappa = kappa_o->typed_data(); out.Cp = Cp_o->typed_data();
  out.tag = tag_o->typed_data();
  return Status(sphb200_advance(&cfg, n, dt, &in, &out, err->typed_data(), *ws, bytes, stream));
}

// neighbor_list.update(position) -> idx[2, capacity], count, err
ffi::Error NeighborsImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                         std::string_view cfg_blob, int32_t mask_self, F32 r, RI32 idx,
                         ffi::Result<ffi::Buffer<ffi::S64>> count, RU32 err) {
  sphb200_config cfg;
  if (auto e = LoadConfig(cfg_blob, &cfg); e.failure()) return e;
  const int64_t n = r.dimensions()[0], cap = idx->dimensions()[1];
  size_t bytes = 0;
  if (int rc = sphb200_workspace_bytes(&cfg, n, &bytes)) return Status(rc);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "sphb200 workspace");
  return Status(sphb200_neighbor_list(&cfg, n, r.typed_data(), idx->typed_data(), cap, mask_self,
                                      count->typed_data(), err->typed_data(), *ws, bytes, stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sphb200_ffi_advance, AdvanceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Attr<std::string_view>("config")
        .Attr<double>("dt")
        .Arg<F32>().Arg<I32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<I32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Ret<ffi::Buffer<ffi::U32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sphb200_ffi_neighbors, NeighborsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Attr<std::string_view>("config")
        .Attr<int32_t>("mask_self")
        .Arg<F32>()
        .Ret<I32>()
        .Ret<ffi::Buffer<ffi::S64>>()
        .Ret<ffi::Buffer<ffi::U32>>());
