// Minimal stand-in for protobuf enum reflection (oracle build only).
#ifndef ORACLE_SHIM_PROTOBUF_DESCRIPTOR_H_
#define ORACLE_SHIM_PROTOBUF_DESCRIPTOR_H_
#include <map>
#include <string>
namespace google { namespace protobuf {
class EnumValueDescriptor {
 public:
  EnumValueDescriptor(const std::string& n, int v) : name_(n), number_(v) {}
  const std::string& name() const { return name_; }
  int number() const { return number_; }
 private:
  std::string name_; int number_;
};
class EnumDescriptor {
 public:
  void Add(const std::string& n, int v) {
    by_number_.insert(std::make_pair(v, new EnumValueDescriptor(n, v)));
    by_name_[n] = v;
  }
  const EnumValueDescriptor* FindValueByNumber(int v) const {
    auto it = by_number_.find(v);
    return it == by_number_.end() ? nullptr : it->second;
  }
  const EnumValueDescriptor* FindValueByName(const std::string& n) const {
    auto it = by_name_.find(n);
    return it == by_name_.end() ? nullptr : FindValueByNumber(it->second);
  }
 private:
  std::map<int, EnumValueDescriptor*> by_number_;
  std::map<std::string, int> by_name_;
};
}}
#endif
