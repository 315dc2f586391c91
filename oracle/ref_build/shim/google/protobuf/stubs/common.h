// Minimal stand-in for protobuf's Mutex helpers (oracle build only).
#ifndef ORACLE_SHIM_PROTOBUF_COMMON_H_
#define ORACLE_SHIM_PROTOBUF_COMMON_H_
#include <mutex>
namespace google { namespace protobuf {
class Mutex {
 public:
  void Lock() { m_.lock(); }
  void Unlock() { m_.unlock(); }
  void AssertHeld() {}
 private:
  std::mutex m_;
};
class MutexLock {
 public:
  explicit MutexLock(Mutex* m) : m_(m) { m_->Lock(); }
  ~MutexLock() { m_->Unlock(); }
 private:
  Mutex* m_;
};
class MutexLockMaybe {
 public:
  explicit MutexLockMaybe(Mutex* m) : m_(m) { if (m_) m_->Lock(); }
  ~MutexLockMaybe() { if (m_) m_->Unlock(); }
 private:
  Mutex* m_;
};
}}
#endif
