// API-SHAPED STUB of the TensorRT 10 plugin headers -- test infrastructure, NOT TensorRT.
//
// This image has no TensorRT, so co-detr-tensorrt_b200/csrc/deformable_attention_plugin_b200.cpp
// cannot be compiled against the real <NvInfer.h> here.  This file declares, with TensorRT 10's names, enum values,
// member order and virtual signatures (as used by the reference's plugin, codetr/csrc/deformable_attention_plugin.cpp,
// and documented for IPluginV3), exactly the types that translation unit touches, so that tests/test_plugin_class.py
// can (1) compile it, (2) drive creator -> plugin -> serialise -> deserialise -> supportsFormatCombination ->
// enqueue through the same virtual calls TensorRT makes.  It proves the file is well-formed C++ against this API
// shape and that its logic is right; it cannot prove ABI compatibility with a real libnvinfer (vtable layout of
// classes this stub leaves out), which needs a TensorRT install.
#pragma once
#include <cstddef>
#include <cstdint>

struct CUstream_st;
typedef CUstream_st *cudaStream_t;

namespace nvinfer1 {

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4, kUINT8 = 5, kFP8 = 6, kBF16 = 7, kINT64 = 8, kINT4 = 9 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3, kCHW16 = 4, kCHW32 = 5 };
enum class PluginFieldType : int32_t {
  kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8, kBF16 = 9, kINT64 = 10, kFP8 = 11, kINT4 = 12
};
enum class PluginCapabilityType : int32_t { kCORE = 0, kBUILD = 1, kRUNTIME = 2 };
enum class TensorRTPhase : int32_t { kBUILD = 0, kRUNTIME = 1 };
enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };

struct Dims {
  static constexpr int32_t MAX_DIMS = 8;
  int32_t nbDims;
  int64_t d[MAX_DIMS];
};
struct PluginTensorDesc {
  Dims dims;
  DataType type;
  TensorFormat format;
  float scale;
};
struct DynamicPluginTensorDesc {
  PluginTensorDesc desc;
  Dims min, max, opt;
};

struct PluginField {
  char const *name;
  void const *data;
  PluginFieldType type;
  int32_t length;
  PluginField(char const *name_ = nullptr, void const *data_ = nullptr, PluginFieldType type_ = PluginFieldType::kUNKNOWN, int32_t length_ = 0) noexcept
      : name(name_), data(data_), type(type_), length(length_) {}
};
struct PluginFieldCollection {
  int32_t nbFields{};
  PluginField const *fields{};
};

class ILogger {
public:
  enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
  virtual void log(Severity severity, char const *msg) noexcept = 0;
  virtual ~ILogger() = default;
};
class ILoggerFinder {
public:
  virtual ILogger *findLogger() = 0;
  virtual ~ILoggerFinder() = default;
};

class IDimensionExpr {
public:
  virtual bool isConstant() const noexcept = 0;
  virtual int64_t getConstantValue() const noexcept = 0;
  virtual ~IDimensionExpr() = default;
};
struct DimsExprs {
  int32_t nbDims;
  IDimensionExpr const *d[Dims::MAX_DIMS];
};
class IExprBuilder {
public:
  virtual IDimensionExpr const *constant(int64_t value) noexcept = 0;
  virtual IDimensionExpr const *operation(DimensionOperation op, IDimensionExpr const &first, IDimensionExpr const &second) noexcept = 0;
  virtual ~IExprBuilder() = default;
};

class IPluginResourceContext {
public:
  virtual ~IPluginResourceContext() = default;
};

class IPluginCapability {
public:
  virtual ~IPluginCapability() = default;
};
class IPluginV3 {
public:
  virtual IPluginCapability *getCapabilityInterface(PluginCapabilityType type) noexcept = 0;
  virtual IPluginV3 *clone() noexcept = 0;
  virtual ~IPluginV3() = default;
};
class IPluginV3OneCore : public IPluginCapability {
public:
  virtual char const *getPluginName() const noexcept = 0;
  virtual char const *getPluginVersion() const noexcept = 0;
  virtual char const *getPluginNamespace() const noexcept = 0;
};
class IPluginV3OneBuild : public IPluginCapability {
public:
  virtual int32_t configurePlugin(DynamicPluginTensorDesc const *in, int32_t nbInputs, DynamicPluginTensorDesc const *out, int32_t nbOutputs) noexcept = 0;
  virtual int32_t getOutputDataTypes(DataType *outputTypes, int32_t nbOutputs, DataType const *inputTypes, int32_t nbInputs) const noexcept = 0;
  virtual int32_t getOutputShapes(DimsExprs const *inputs, int32_t nbInputs, DimsExprs const *shapeInputs, int32_t nbShapeInputs, DimsExprs *outputs,
                                  int32_t nbOutputs, IExprBuilder &exprBuilder) noexcept = 0;
  virtual bool supportsFormatCombination(int32_t pos, DynamicPluginTensorDesc const *inOut, int32_t nbInputs, int32_t nbOutputs) noexcept = 0;
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual size_t getWorkspaceSize(DynamicPluginTensorDesc const *, int32_t, DynamicPluginTensorDesc const *, int32_t) const noexcept { return 0; }
  virtual int32_t getValidTactics(int32_t *, int32_t) noexcept { return 0; }
  virtual int32_t getNbTactics() noexcept { return 0; }
  virtual char const *getTimingCacheID() noexcept { return nullptr; }
  virtual int32_t getFormatCombinationLimit() noexcept { return 100; }
  virtual char const *getMetadataString() noexcept { return nullptr; }
};
class IPluginV3OneRuntime : public IPluginCapability {
public:
  virtual int32_t setTactic(int32_t) noexcept { return 0; }
  virtual int32_t onShapeChange(PluginTensorDesc const *in, int32_t nbInputs, PluginTensorDesc const *out, int32_t nbOutputs) noexcept = 0;
  virtual int32_t enqueue(PluginTensorDesc const *inputDesc, PluginTensorDesc const *outputDesc, void const *const *inputs, void *const *outputs,
                          void *workspace, cudaStream_t stream) noexcept = 0;
  virtual IPluginV3 *attachToContext(IPluginResourceContext *context) noexcept = 0;
  virtual PluginFieldCollection const *getFieldsToSerialize() noexcept = 0;
};

class IPluginCreatorInterface {
public:
  virtual ~IPluginCreatorInterface() = default;
};
class IPluginCreatorV3One : public IPluginCreatorInterface {
public:
  virtual IPluginV3 *createPlugin(char const *name, PluginFieldCollection const *fc, TensorRTPhase phase) noexcept = 0;
  virtual PluginFieldCollection const *getFieldNames() noexcept = 0;
  virtual char const *getPluginName() const noexcept = 0;
  virtual char const *getPluginVersion() const noexcept = 0;
  virtual char const *getPluginNamespace() const noexcept = 0;
};

// stub registry: REGISTER_TENSORRT_PLUGIN(T) constructs one static T and records it (the harness reads the list back)
struct StubRegistry {
  static constexpr int kMax = 8;
  IPluginCreatorInterface *creators[kMax];
  int count;
  static StubRegistry &get() {
    static StubRegistry r{};
    return r;
  }
  void add(IPluginCreatorInterface *c) {
    if (count < kMax) creators[count++] = c;
  }
};
template <typename T>
class PluginRegistrar {
public:
  PluginRegistrar() { StubRegistry::get().add(&instance); }

private:
  T instance{};
};

}  // namespace nvinfer1

nvinfer1::ILogger *getLogger() noexcept;  // the stub's: defined by the harness

#define REGISTER_TENSORRT_PLUGIN(name) static nvinfer1::PluginRegistrar<name> pluginRegistrar##name {}
