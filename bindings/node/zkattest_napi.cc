// zkattest_napi.cc — node-addon-api shim over the C ABI (include/zkattest.h).
// NOT compiled in this image (no node / node-addon-api headers).  Build where node >= 24 exists:
//   npm i node-addon-api && node-gyp configure build   (binding.gyp links -lzkattest)
// Every call is a Napi::AsyncWorker so the TypeScript functions keep returning Promises
// (the reference is async only because of WebCrypto, src/zkpAttestList.ts:104,147).
#include <napi.h>

#include <string>
#include <vector>

#include "zkattest.h"

namespace {

zka_ctx* g_ctx = nullptr;

const char* status_message(int s) {
  switch (s) {
    case ZKA_ERR_INVALID_PK: return "invalid public key";          // zkpAttestList.ts:117
    case ZKA_ERR_T_INFINITY: return "T[i] is at infinity";         // exp.ts:151
    case ZKA_ERR_T1_INFINITY: return "T1 is at infinity";          // exp.ts:193
    case ZKA_ERR_POINTS_DONT_ADD: return "Points don't add up!";   // pointAdd.ts:105
    case ZKA_ERR_R_INFINITY: return "R is at infinity";            // zkpAttestList.ts:159
    case ZKA_ERR_MALFORMED: return "error deserializing Point";    // weier.ts:87
    case ZKA_ERR_PARAMS_NOT_FOUND: return "params not found";      // exp.ts:270
    default: return "zkattest: internal status";
  }
}

struct Params : public Napi::ObjectWrap<Params> {
  zka_params* h = nullptr;
  static Napi::Function Init(Napi::Env env) { return DefineClass(env, "Params", {}); }
  explicit Params(const Napi::CallbackInfo& info) : Napi::ObjectWrap<Params>(info) {}
  ~Params() { zka_params_destroy(h); }
};

// proveBatch(params, msgHash: Uint8Array[B*32], sig[B*64], pk[B*65], which: Uint32Array[B],
//            ring: Uint8Array[N*32], tape: Uint8Array[B*stride]) -> Promise<{proofs, lens, status}>
class ProveWorker : public Napi::AsyncWorker {
 public:
  ProveWorker(Napi::Env env, zka_params* p, std::vector<uint8_t> msg, std::vector<uint8_t> sig,
              std::vector<uint8_t> pk, std::vector<uint32_t> which, std::vector<uint8_t> ring,
              std::vector<uint8_t> tape, uint32_t sec)
      : Napi::AsyncWorker(env), deferred(Napi::Promise::Deferred::New(env)), p_(p), msg_(std::move(msg)),
        sig_(std::move(sig)), pk_(std::move(pk)), which_(std::move(which)), ring_(std::move(ring)),
        tape_(std::move(tape)), sec_(sec) {}
  void Execute() override {
    const uint32_t B = (uint32_t)which_.size(), N = (uint32_t)(ring_.size() / 32);
    stride_ = zka_proof_max_len(N, sec_);
    proofs_.resize((size_t)B * stride_);
    lens_.resize(B);
    status_.resize(B);
    int rc = zka_prove_batch(g_ctx, p_, B, msg_.data(), sig_.data(), pk_.data(), which_.data(), ring_.data(), N,
                             tape_.data(), tape_.size() / B, proofs_.data(), stride_, lens_.data(), status_.data());
    if (rc != 0) SetError(zka_last_error(g_ctx));
  }
  void OnOK() override {
    Napi::Env env = Env();
    for (size_t i = 0; i < status_.size(); i++)
      if (status_[i] != 0) { deferred.Reject(Napi::Error::New(env, status_message(status_[i])).Value()); return; }
    Napi::Object o = Napi::Object::New(env);
    o.Set("proofs", Napi::Buffer<uint8_t>::Copy(env, proofs_.data(), proofs_.size()));
    o.Set("stride", Napi::Number::New(env, (double)stride_));
    o.Set("lens", Napi::Buffer<uint32_t>::Copy(env, lens_.data(), lens_.size()));
    deferred.Resolve(o);
  }
  void OnError(const Napi::Error& e) override { deferred.Reject(e.Value()); }
  Napi::Promise::Deferred deferred;

 private:
  zka_params* p_;
  std::vector<uint8_t> msg_, sig_, pk_;
  std::vector<uint32_t> which_;
  std::vector<uint8_t> ring_, tape_, proofs_;
  std::vector<uint32_t> lens_;
  std::vector<int32_t> status_;
  uint32_t sec_;
  size_t stride_ = 0;
};

template <class T>
std::vector<T> to_vec(const Napi::Value& v) {
  auto a = v.As<Napi::TypedArrayOf<T>>();
  return std::vector<T>(a.Data(), a.Data() + a.ElementLength());
}

Napi::Value ProveBatch(const Napi::CallbackInfo& info) {
  Params* P = Napi::ObjectWrap<Params>::Unwrap(info[0].As<Napi::Object>());
  auto* w = new ProveWorker(info.Env(), P->h, to_vec<uint8_t>(info[1]), to_vec<uint8_t>(info[2]),
                            to_vec<uint8_t>(info[3]), to_vec<uint32_t>(info[4]), to_vec<uint8_t>(info[5]),
                            to_vec<uint8_t>(info[6]), info[7].As<Napi::Number>().Uint32Value());
  w->Queue();
  return w->deferred.Promise();
}

// paramsGenerate(rnd64) -> {hNist: Buffer(65), hProof: Buffer(67)};  paramsCreate(hNist, hProof, secLevel) -> Params
Napi::Value ParamsGenerate(const Napi::CallbackInfo& info) {
  auto rnd = info[0].As<Napi::Uint8Array>();
  uint8_t hn[65], hp[67];
  if (zka_params_generate(g_ctx, rnd.Data(), hn, hp) != 0)
    Napi::Error::New(info.Env(), zka_last_error(g_ctx)).ThrowAsJavaScriptException();
  Napi::Object o = Napi::Object::New(info.Env());
  o.Set("hNist", Napi::Buffer<uint8_t>::Copy(info.Env(), hn, 65));
  o.Set("hProof", Napi::Buffer<uint8_t>::Copy(info.Env(), hp, 67));
  return o;
}

Napi::Object InitAll(Napi::Env env, Napi::Object exports) {
  if (zka_init(0, &g_ctx) != 0) Napi::Error::New(env, "zkattest: no CUDA device (no CPU fallback)").ThrowAsJavaScriptException();
  exports.Set("Params", Params::Init(env));
  exports.Set("paramsGenerate", Napi::Function::New(env, ParamsGenerate));
  exports.Set("proveBatch", Napi::Function::New(env, ProveBatch));
  // verifyBatch / keyToInt / paramsCreate follow the same pattern over zka_verify_batch,
  // zka_key_to_int and zka_params_create.
  return exports;
}

}  // namespace

NODE_API_MODULE(zkattest, InitAll)
