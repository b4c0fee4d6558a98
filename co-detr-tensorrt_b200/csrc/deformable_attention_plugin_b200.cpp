// deformable_attention_plugin_b200.cpp -- TensorRT IPluginV3 for multi-scale deformable attention WITHOUT libtorch.
//
// Drop-in for the reference's libdeformable_attention_plugin.so (codetr/csrc/deformable_attention_plugin.cpp): same
// plugin identity ("DeformableAttentionPlugin", version "1", namespace "", :77-79), same creator field (`im2col_step`,
// kINT64, length 1, :398), same serialised field ("parameters", kUNKNOWN, the raw 8-byte struct {int64 im2col_step},
// :84-86 / :381-388), same I/O contract (5 inputs / 1 output, kLINEAR, value / sampling_loc / attn_weight / output of
// one floating type, spatial_shapes / level_start_index kINT64, :218-246), same exported C symbols
// (`getPluginCreators`, `setLoggerFinder`, :507-514) and the static registration (:466) -- so an engine built with the
// reference's plugin deserialises against this library and `ops.py`'s dynamo converter (:189-291) finds the creator
// under the same name.
//
// What differs, deliberately:
//   * enqueue (:285-355) hands TensorRT's raw device pointers and stream straight to the C ABI
//     (msda_b200_plugin_enqueue, include/msda_b200.h).  No at::from_blob, no stream guard, no libtorch / libc10 in the
//     engine runtime's link line (the reference links all of ATen for six tensor wrappers).
//   * kBF16 is accepted next to kFLOAT / kHALF in supportsFormatCombination.
//   * configurePlugin reports a malformed network with a logged error and a non-zero return instead of abort().
// The plugin is stateless apart from im2col_step: clone() / attachToContext() copy one integer, enqueue is re-entrant
// and capture-safe (the C ABI never allocates, synchronises or reads device memory on the host).
//
// Build (needs the TensorRT headers; this image has none -- tests/test_plugin_class.py compiles it against the
// API-shaped stub headers under tests/stubs/):
//   g++ -std=c++17 -O2 -fPIC -shared -I<TensorRT>/include -I/usr/local/cuda/include -Iinclude
//       co-detr-tensorrt_b200/csrc/deformable_attention_plugin_b200.cpp -Lco-detr-tensorrt_b200/csrc -lmsda_b200
//       -Wl,-rpath,'$ORIGIN' -lnvinfer -o libdeformable_attention_plugin.so
#include <NvInfer.h>
#include <NvInferPlugin.h>
#include <NvInferRuntime.h>
#include <NvInferRuntimePlugin.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "msda_b200.h"

namespace codetr_b200 {

using namespace nvinfer1;

constexpr char const *kName = "DeformableAttentionPlugin";
constexpr char const *kVersion = "1";
constexpr char const *kNamespace = "";
constexpr int32_t kNumInputs = 5;   // value, spatial_shapes, level_start_index, sampling_loc, attn_weight
constexpr int32_t kNumOutputs = 1;  // [bs, num_queries, num_heads * dim_per_head]

// The serialised state, bit-compatible with the reference's DeformableAttentionParameters (plugin.cpp:84-86).
struct Parameters {
  int64_t im2col_step;
};
static_assert(sizeof(Parameters) == 8, "engines serialised by the reference plugin carry exactly 8 bytes");

// ---- logging: TensorRT's logger when the library was loaded through the plugin registry, stderr otherwise ----
class LoggerSlot {
public:
  void set(ILoggerFinder *finder) {
    std::lock_guard<std::mutex> lock(mu_);
    if (!finder_ && finder) finder_ = finder;
  }
  void error(std::string const &msg) {
    ILogger *logger = nullptr;
    {
      std::lock_guard<std::mutex> lock(mu_);
      if (finder_) logger = finder_->findLogger();
    }
    if (!logger) logger = getLogger();
    if (logger) logger->log(ILogger::Severity::kERROR, msg.c_str());
    else std::fprintf(stderr, "[DeformableAttentionPlugin] %s\n", msg.c_str());
  }

private:
  std::mutex mu_;
  ILoggerFinder *finder_ = nullptr;
};
LoggerSlot g_log;

// true when `ok`; otherwise logs what was expected and returns false (never aborts the host process)
bool expect(bool ok, char const *what) {
  if (!ok) g_log.error(std::string("validation failed: ") + what);
  return ok;
}

bool is_value_type(DataType t) { return t == DataType::kFLOAT || t == DataType::kHALF || t == DataType::kBF16; }

class MsdaPluginV3 final : public IPluginV3, public IPluginV3OneCore, public IPluginV3OneBuild, public IPluginV3OneRuntime {
public:
  explicit MsdaPluginV3(Parameters const &params) : params_(params) {
    field_ = PluginField("parameters", &params_, PluginFieldType::kUNKNOWN, static_cast<int32_t>(sizeof(Parameters)));
    fields_.nbFields = 1;
    fields_.fields = &field_;
  }
  MsdaPluginV3(MsdaPluginV3 const &) = delete;
  MsdaPluginV3 &operator=(MsdaPluginV3 const &) = delete;

  // ---- IPluginV3 ----
  IPluginCapability *getCapabilityInterface(PluginCapabilityType type) noexcept override {
    switch (type) {
    case PluginCapabilityType::kBUILD: return static_cast<IPluginV3OneBuild *>(this);
    case PluginCapabilityType::kRUNTIME: return static_cast<IPluginV3OneRuntime *>(this);
    case PluginCapabilityType::kCORE: return static_cast<IPluginV3OneCore *>(this);
    }
    return nullptr;
  }
  IPluginV3 *clone() noexcept override { return new (std::nothrow) MsdaPluginV3(params_); }

  // ---- IPluginV3OneCore ----
  char const *getPluginName() const noexcept override { return kName; }
  char const *getPluginVersion() const noexcept override { return kVersion; }
  char const *getPluginNamespace() const noexcept override { return kNamespace; }

  // ---- IPluginV3OneBuild ----
  int32_t getNbOutputs() const noexcept override { return kNumOutputs; }

  // Rank / extent agreement of the five inputs and the output (plugin.cpp:151-216 asserts the same relations).
  int32_t configurePlugin(DynamicPluginTensorDesc const *in, int32_t nbInputs, DynamicPluginTensorDesc const *out,
                          int32_t nbOutputs) noexcept override {
    if (!expect(nbInputs == kNumInputs && nbOutputs == kNumOutputs && in && out, "5 inputs, 1 output")) return 1;
    Dims const &v = in[0].desc.dims, &shp = in[1].desc.dims, &st = in[2].desc.dims, &loc = in[3].desc.dims, &w = in[4].desc.dims,
               &o = out[0].desc.dims;
    bool ok = expect(v.nbDims == 4, "value is [bs, num_keys, num_heads, dim_per_head]") &&
              expect(shp.nbDims == 2, "spatial_shapes is [num_levels, 2]") && expect(st.nbDims == 1, "level_start_index is [num_levels]") &&
              expect(loc.nbDims == 6, "sampling_loc is [bs, num_queries, num_heads, num_levels, num_points, 2]") &&
              expect(w.nbDims == 5, "attn_weight is [bs, num_queries, num_heads, num_levels, num_points]") &&
              expect(o.nbDims == 3, "output is [bs, num_queries, num_heads * dim_per_head]");
    if (!ok) return 1;
    ok = expect(v.d[0] == loc.d[0] && v.d[0] == w.d[0] && v.d[0] == o.d[0], "one batch size") &&
         expect(loc.d[1] == w.d[1] && loc.d[1] == o.d[1], "one query count") &&
         expect(v.d[2] == loc.d[2] && v.d[2] == w.d[2], "one head count") &&
         expect(shp.d[0] == st.d[0] && shp.d[0] == loc.d[3] && shp.d[0] == w.d[3], "one level count") &&
         expect(loc.d[4] == w.d[4], "one point count") && expect(o.d[2] == v.d[2] * v.d[3], "output width = num_heads * dim_per_head");
    return ok ? 0 : 1;
  }

  // pos 0, 3, 4, 5 (value, sampling_loc, attn_weight, output): one of fp32 / fp16 / bf16, all equal to input 0's type;
  // pos 1, 2 (spatial_shapes, level_start_index): int64; everything linear (plugin.cpp:218-246, plus kBF16).
  bool supportsFormatCombination(int32_t pos, DynamicPluginTensorDesc const *inOut, int32_t nbInputs, int32_t nbOutputs) noexcept override {
    if (nbInputs != kNumInputs || nbOutputs != kNumOutputs || pos < 0 || pos >= kNumInputs + kNumOutputs || !inOut) return false;
    PluginTensorDesc const &d = inOut[pos].desc;
    if (d.format != TensorFormat::kLINEAR) return false;
    if (pos == 1 || pos == 2) return d.type == DataType::kINT64;
    return is_value_type(d.type) && d.type == inOut[0].desc.type;
  }

  int32_t getOutputDataTypes(DataType *outputTypes, int32_t nbOutputs, DataType const *inputTypes, int32_t nbInputs) const noexcept override {
    if (nbInputs != kNumInputs || nbOutputs != kNumOutputs || !outputTypes || !inputTypes) return 1;
    outputTypes[0] = inputTypes[0];
    return 0;
  }

  // output = [value.d0, sampling_loc.d1, value.d2 * value.d3]   (plugin.cpp:257-281)
  int32_t getOutputShapes(DimsExprs const *inputs, int32_t nbInputs, DimsExprs const *, int32_t, DimsExprs *outputs, int32_t nbOutputs,
                          IExprBuilder &exprBuilder) noexcept override {
    if (nbInputs != kNumInputs || nbOutputs != kNumOutputs || !inputs || !outputs || inputs[0].nbDims != 4 || inputs[3].nbDims != 6) return 1;
    outputs[0].nbDims = 3;
    outputs[0].d[0] = inputs[0].d[0];
    outputs[0].d[1] = inputs[3].d[1];
    outputs[0].d[2] = exprBuilder.operation(DimensionOperation::kPROD, *inputs[0].d[2], *inputs[0].d[3]);
    return 0;
  }

  // The default kernels gather from the op's own layouts and need no scratch memory (the reference: 0, plugin.cpp:371).
  // msda_b200_plugin_workspace_bytes() reports what the opt-in packed-pyramid path would take; it measured slower
  // than the default path, so the engine is not asked to reserve it.
  size_t getWorkspaceSize(DynamicPluginTensorDesc const *, int32_t, DynamicPluginTensorDesc const *, int32_t) const noexcept override { return 0; }

  // ---- IPluginV3OneRuntime ----
  int32_t onShapeChange(PluginTensorDesc const *, int32_t, PluginTensorDesc const *, int32_t) noexcept override { return 0; }

  // Raw device pointers + TensorRT's stream -> C ABI.  Returns non-zero on failure like the reference (:320-325); the
  // message goes to the logger.
  int32_t enqueue(PluginTensorDesc const *inputDesc, PluginTensorDesc const *outputDesc, void const *const *inputs, void *const *outputs,
                  void *workspace, cudaStream_t stream) noexcept override {
    (void)outputDesc;
    if (!inputDesc || !inputs || !outputs) return 1;
    Dims const &v = inputDesc[0].dims, &loc = inputDesc[3].dims;
    if (!expect(v.nbDims == 4 && loc.nbDims == 6, "enqueue: value rank 4, sampling_loc rank 6")) return 1;
    int64_t value_dims[4], loc_dims[6];
    for (int i = 0; i < 4; ++i) value_dims[i] = v.d[i];
    for (int i = 0; i < 6; ++i) loc_dims[i] = loc.d[i];
    // workspace size 0 was requested: whatever TensorRT passes is not used
    int const rc = msda_b200_plugin_enqueue(value_dims, loc_dims, static_cast<int>(inputDesc[0].type), inputs, outputs, nullptr, 0,
                                            params_.im2col_step, stream);
    (void)workspace;
    if (rc != 0) {
      g_log.error(std::string("enqueue failed: ") + msda_b200_error_string(rc));
      return 1;
    }
    return 0;
  }

  IPluginV3 *attachToContext(IPluginResourceContext *) noexcept override { return clone(); }
  PluginFieldCollection const *getFieldsToSerialize() noexcept override { return &fields_; }

private:
  Parameters params_;
  PluginField field_;
  PluginFieldCollection fields_;
};

class MsdaPluginCreator final : public IPluginCreatorV3One {
public:
  MsdaPluginCreator() {
    attribute_ = PluginField("im2col_step", nullptr, PluginFieldType::kINT64, 1);
    attributes_.nbFields = 1;
    attributes_.fields = &attribute_;
  }

  PluginFieldCollection const *getFieldNames() noexcept override { return &attributes_; }

  // kBUILD: `im2col_step` (kINT64) from the network definition, default 64 like the reference (:417).
  // kRUNTIME: the single serialised field written by getFieldsToSerialize() -- by this plugin or by the reference's.
  IPluginV3 *createPlugin(char const *, PluginFieldCollection const *fc, TensorRTPhase phase) noexcept override {
    Parameters params{64};
    if (phase == TensorRTPhase::kBUILD) {
      for (int32_t i = 0; fc && i < fc->nbFields; ++i) {
        PluginField const &f = fc->fields[i];
        if (f.name && std::strcmp(f.name, "im2col_step") == 0 && f.type == PluginFieldType::kINT64 && f.data) {
          std::memcpy(&params.im2col_step, f.data, sizeof(int64_t));
        }
      }
    } else if (phase == TensorRTPhase::kRUNTIME) {
      if (!expect(fc && fc->nbFields == 1 && fc->fields, "one serialised field")) return nullptr;
      PluginField const &f = fc->fields[0];
      if (!expect(f.name && std::strcmp(f.name, "parameters") == 0 && f.type == PluginFieldType::kUNKNOWN &&
                      f.length == static_cast<int32_t>(sizeof(Parameters)) && f.data,
                  "field `parameters`: kUNKNOWN, 8 bytes"))
        return nullptr;
      std::memcpy(&params, f.data, sizeof(Parameters));
    } else {
      return nullptr;
    }
    if (!expect(params.im2col_step > 0, "im2col_step > 0")) return nullptr;
    return new (std::nothrow) MsdaPluginV3(params);
  }

  char const *getPluginName() const noexcept override { return kName; }
  char const *getPluginVersion() const noexcept override { return kVersion; }
  char const *getPluginNamespace() const noexcept override { return kNamespace; }

private:
  PluginField attribute_;
  PluginFieldCollection attributes_;
};

// found by `registry.get_creator("DeformableAttentionPlugin", "1")` as soon as the library is dlopen'ed (plugin.cpp:466)
REGISTER_TENSORRT_PLUGIN(MsdaPluginCreator);

}  // namespace codetr_b200

// ---- the two C entry points TensorRT's plugin-library loader looks for (plugin.cpp:507-514) ----
extern "C" void setLoggerFinder(nvinfer1::ILoggerFinder *finder) { codetr_b200::g_log.set(finder); }

extern "C" nvinfer1::IPluginCreatorInterface *const *getPluginCreators(int32_t &nbCreators) {
  static codetr_b200::MsdaPluginCreator creator;
  static nvinfer1::IPluginCreatorInterface *const list[] = {&creator};
  nbCreators = 1;
  return list;
}
