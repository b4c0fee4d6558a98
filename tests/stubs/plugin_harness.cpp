// Test harness (tests/test_plugin_class.py): drives deformable_attention_plugin_b200.cpp through the virtual calls
// TensorRT makes, against the stub headers in this directory.  Plain C entry points for ctypes.
#include <cstdio>
#include <cstring>

#include "NvInfer.h"

extern "C" nvinfer1::IPluginCreatorInterface *const *getPluginCreators(int32_t &nbCreators);
extern "C" void setLoggerFinder(nvinfer1::ILoggerFinder *finder);

namespace {
using namespace nvinfer1;

struct CountingLogger : ILogger {
  int errors = 0;
  char last[512] = {0};
  void log(Severity severity, char const *msg) noexcept override {
    if (severity <= Severity::kERROR) {
      ++errors;
      std::snprintf(last, sizeof(last), "%s", msg);
    }
  }
};
CountingLogger g_logger;
struct Finder : ILoggerFinder {
  ILogger *findLogger() override { return &g_logger; }
} g_finder;

struct ConstExpr : IDimensionExpr {
  int64_t v;
  explicit ConstExpr(int64_t value = 0) : v(value) {}
  bool isConstant() const noexcept override { return true; }
  int64_t getConstantValue() const noexcept override { return v; }
};
struct Builder : IExprBuilder {
  ConstExpr pool[64];
  int used = 0;
  IDimensionExpr const *constant(int64_t value) noexcept override {
    pool[used] = ConstExpr(value);
    return &pool[used++];
  }
  IDimensionExpr const *operation(DimensionOperation op, IDimensionExpr const &a, IDimensionExpr const &b) noexcept override {
    int64_t const x = a.getConstantValue(), y = b.getConstantValue();
    return constant(op == DimensionOperation::kPROD ? x * y : op == DimensionOperation::kSUM ? x + y : 0);
  }
};

IPluginCreatorV3One *creator() {
  int32_t n = 0;
  auto list = getPluginCreators(n);
  return n == 1 ? static_cast<IPluginCreatorV3One *>(list[0]) : nullptr;
}
IPluginV3 *g_plugin = nullptr;

void fill_descs(DynamicPluginTensorDesc *io, int64_t const *vd, int64_t const *ld, int dtype) {
  std::memset(io, 0, sizeof(DynamicPluginTensorDesc) * 6);
  auto set = [&](int i, int nb, int64_t const *d, DataType t) {
    io[i].desc.dims.nbDims = nb;
    for (int k = 0; k < nb; ++k) io[i].desc.dims.d[k] = d[k];
    io[i].desc.type = t;
    io[i].desc.format = TensorFormat::kLINEAR;
  };
  int64_t const shp[2] = {ld[3], 2}, st[1] = {ld[3]}, w[5] = {ld[0], ld[1], ld[2], ld[3], ld[4]}, o[3] = {vd[0], ld[1], vd[2] * vd[3]};
  set(0, 4, vd, static_cast<DataType>(dtype));
  set(1, 2, shp, DataType::kINT64);
  set(2, 1, st, DataType::kINT64);
  set(3, 6, ld, static_cast<DataType>(dtype));
  set(4, 5, w, static_cast<DataType>(dtype));
  set(5, 3, o, static_cast<DataType>(dtype));
}
}  // namespace

nvinfer1::ILogger *getLogger() noexcept { return &g_logger; }

extern "C" {

int harness_registered_creators() { return nvinfer1::StubRegistry::get().count; }
int harness_logged_errors() { return g_logger.errors; }
char const *harness_last_error() { return g_logger.last; }
void harness_use_logger_finder() { setLoggerFinder(&g_finder); }

// identity of the creator and its build-phase field list: name|version|namespace|field name|field type|field length
int harness_creator_identity(char *out, int cap) {
  auto *c = creator();
  if (!c) return 1;
  auto const *fc = c->getFieldNames();
  std::snprintf(out, cap, "%s|%s|%s|%s|%d|%d|%d", c->getPluginName(), c->getPluginVersion(), c->getPluginNamespace(), fc->fields[0].name,
                static_cast<int>(fc->fields[0].type), fc->fields[0].length, fc->nbFields);
  return 0;
}

// build-phase creation with the `im2col_step` field (or with no field when step < 0); keeps the plugin for later calls
int harness_create_build(long long step) {
  auto *c = creator();
  if (!c) return 1;
  int64_t value = step;
  PluginField f("im2col_step", &value, PluginFieldType::kINT64, 1);
  PluginFieldCollection fc;
  fc.nbFields = step < 0 ? 0 : 1;
  fc.fields = &f;
  delete g_plugin;
  g_plugin = c->createPlugin("msda", &fc, TensorRTPhase::kBUILD);
  return g_plugin ? 0 : 2;
}

// the serialised field of the current plugin: name|type|length|nbFields and the raw bytes
int harness_serialised(char *desc, int cap, unsigned char *bytes, int bytes_cap) {
  if (!g_plugin) return 1;
  auto *rt = static_cast<IPluginV3OneRuntime *>(g_plugin->getCapabilityInterface(PluginCapabilityType::kRUNTIME));
  auto const *fc = rt->getFieldsToSerialize();
  std::snprintf(desc, cap, "%s|%d|%d|%d", fc->fields[0].name, static_cast<int>(fc->fields[0].type), fc->fields[0].length, fc->nbFields);
  if (fc->fields[0].length > bytes_cap) return 2;
  std::memcpy(bytes, fc->fields[0].data, fc->fields[0].length);
  return 0;
}

// runtime-phase creation from raw serialised bytes (what TensorRT hands back when an engine is deserialised)
int harness_create_runtime(char const *field_name, int field_type, unsigned char const *bytes, int length) {
  auto *c = creator();
  if (!c) return 1;
  PluginField f(field_name, bytes, static_cast<PluginFieldType>(field_type), length);
  PluginFieldCollection fc;
  fc.nbFields = 1;
  fc.fields = &f;
  IPluginV3 *p = c->createPlugin("msda", &fc, TensorRTPhase::kRUNTIME);
  if (!p) return 2;
  delete g_plugin;
  g_plugin = p;
  return 0;
}

int harness_core_identity(char *out, int cap) {
  if (!g_plugin) return 1;
  auto *core = static_cast<IPluginV3OneCore *>(g_plugin->getCapabilityInterface(PluginCapabilityType::kCORE));
  std::snprintf(out, cap, "%s|%s|%s", core->getPluginName(), core->getPluginVersion(), core->getPluginNamespace());
  return 0;
}

// supportsFormatCombination(pos) for the six tensors with the given types / formats
int harness_supports(int pos, int const *types, int const *formats) {
  if (!g_plugin) return -1;
  auto *b = static_cast<IPluginV3OneBuild *>(g_plugin->getCapabilityInterface(PluginCapabilityType::kBUILD));
  DynamicPluginTensorDesc io[6];
  std::memset(io, 0, sizeof(io));
  for (int i = 0; i < 6; ++i) {
    io[i].desc.type = static_cast<DataType>(types[i]);
    io[i].desc.format = static_cast<TensorFormat>(formats[i]);
  }
  return b->supportsFormatCombination(pos, io, 5, 1) ? 1 : 0;
}

// configurePlugin + getOutputShapes + getOutputDataTypes + getWorkspaceSize + getNbOutputs on consistent descriptors;
// out = {configure rc, out d0, d1, d2, out dtype, workspace bytes, nb outputs}
int harness_build_queries(long long const *value_dims, long long const *loc_dims, int dtype, long long *out, int break_heads) {
  if (!g_plugin) return 1;
  auto *b = static_cast<IPluginV3OneBuild *>(g_plugin->getCapabilityInterface(PluginCapabilityType::kBUILD));
  DynamicPluginTensorDesc io[6];
  int64_t vd[4], ld[6];
  for (int i = 0; i < 4; ++i) vd[i] = value_dims[i];
  for (int i = 0; i < 6; ++i) ld[i] = loc_dims[i];
  fill_descs(io, vd, ld, dtype);
  if (break_heads) io[4].desc.dims.d[2] += 1;
  out[0] = b->configurePlugin(io, 5, io + 5, 1);
  Builder eb;
  DimsExprs ins[5];
  for (int i = 0; i < 5; ++i) {
    ins[i].nbDims = io[i].desc.dims.nbDims;
    for (int k = 0; k < ins[i].nbDims; ++k) ins[i].d[k] = eb.constant(io[i].desc.dims.d[k]);
  }
  DimsExprs outs[1];
  if (b->getOutputShapes(ins, 5, nullptr, 0, outs, 1, eb) != 0) return 2;
  for (int k = 0; k < 3; ++k) out[1 + k] = outs[0].d[k]->getConstantValue();
  DataType in_types[5] = {static_cast<DataType>(dtype), DataType::kINT64, DataType::kINT64, static_cast<DataType>(dtype), static_cast<DataType>(dtype)};
  DataType out_type;
  if (b->getOutputDataTypes(&out_type, 1, in_types, 5) != 0) return 3;
  out[4] = static_cast<long long>(out_type);
  out[5] = static_cast<long long>(b->getWorkspaceSize(io, 5, io + 5, 1));
  out[6] = b->getNbOutputs();
  return 0;
}

// enqueue on a context clone (attachToContext), raw pointers and an external stream -- what an execution context does
int harness_enqueue(long long const *value_dims, long long const *loc_dims, int dtype, void const *const *inputs, void *output, void *stream) {
  if (!g_plugin) return -1;
  auto *rt0 = static_cast<IPluginV3OneRuntime *>(g_plugin->getCapabilityInterface(PluginCapabilityType::kRUNTIME));
  IPluginV3 *ctx = rt0->attachToContext(nullptr);
  if (!ctx) return -2;
  auto *rt = static_cast<IPluginV3OneRuntime *>(ctx->getCapabilityInterface(PluginCapabilityType::kRUNTIME));
  DynamicPluginTensorDesc io[6];
  int64_t vd[4], ld[6];
  for (int i = 0; i < 4; ++i) vd[i] = value_dims[i];
  for (int i = 0; i < 6; ++i) ld[i] = loc_dims[i];
  fill_descs(io, vd, ld, dtype);
  PluginTensorDesc in[5], out[1];
  for (int i = 0; i < 5; ++i) in[i] = io[i].desc;
  out[0] = io[5].desc;
  void *outs[1] = {output};
  rt->onShapeChange(in, 5, out, 1);
  int const rc = rt->enqueue(in, out, inputs, outs, nullptr, static_cast<cudaStream_t>(stream));
  delete ctx;
  return rc;
}

}  // extern "C"
