// projector.cc -- column projections (see include/supersonic/projector.h).
#include "supersonic/projector.h"

#include <stdio.h>

namespace supersonic {

bool BoundSingleSourceProjector::AddAs(int source_position, const StringPiece& alias) {
  const Attribute& a = source_schema_.attribute(source_position);
  if (!result_schema_.add_attribute(Attribute(alias.as_string(), a.type(), a.nullability()))) return false;
  positions_.push_back(source_position);
  return true;
}
void BoundSingleSourceProjector::Project(const View& source, View* target) const {
  for (size_t i = 0; i < positions_.size(); ++i) {
    target->mutable_column(static_cast<int>(i))->ResetFrom(source.column(positions_[i]));
  }
  target->set_row_count(source.row_count());
}
// view_copier.h:95-97: copies the projected columns
ViewCopier::ViewCopier(const BoundSingleSourceProjector* projector, bool deep_copy)
    : schema_(projector->result_schema()), deep_copy_(deep_copy) {
  for (int i = 0; i < schema_.attribute_count(); ++i) source_.push_back(projector->source_attribute_position(i));
}

bool BoundSingleSourceProjector::IsAttributeProjected(int source_position) const {
  for (size_t i = 0; i < positions_.size(); ++i) if (positions_[i] == source_position) return true;
  return false;
}

namespace {

Exception* Missing(const string& name, const TupleSchema& schema) {
  return new Exception(ERROR_ATTRIBUTE_MISSING,
                       "No attribute '" + name + "' in the schema: (" + schema.GetHumanReadableSpecification() + ")");
}
Exception* Duplicate(const string& name) {
  return new Exception(ERROR_ATTRIBUTE_EXISTS, "Duplicate attribute name '" + name + "' in result schema");
}

class NamedProjector : public SingleSourceProjector {
 public:
  NamedProjector(const string& name, const string& alias, bool has_alias) : name_(name), alias_(alias), has_alias_(has_alias) {}
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& s) const {
    const int pos = s.LookupAttributePosition(name_);
    if (pos < 0) THROW(Missing(name_, s));
    std::unique_ptr<BoundSingleSourceProjector> b(new BoundSingleSourceProjector(s));
    if (!b->AddAs(pos, has_alias_ ? alias_ : name_)) THROW(Duplicate(alias_));
    return Success(static_cast<const BoundSingleSourceProjector*>(b.release()));
  }
  virtual SingleSourceProjector* Clone() const { return new NamedProjector(name_, alias_, has_alias_); }
  virtual string ToString(bool) const { return has_alias_ ? name_ + " AS " + alias_ : name_; }
 private:
  string name_, alias_;
  bool has_alias_;
};

class PositionProjector : public SingleSourceProjector {
 public:
  PositionProjector(const vector<int>& positions, const vector<string>& aliases) : positions_(positions), aliases_(aliases) {}
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& s) const {
    std::unique_ptr<BoundSingleSourceProjector> b(new BoundSingleSourceProjector(s));
    for (size_t i = 0; i < positions_.size(); ++i) {
      if (positions_[i] < 0 || positions_[i] >= s.attribute_count()) {
        char buf[160];
        snprintf(buf, sizeof(buf), "Attribute position %d out of range; the schema has %d attributes",
                 positions_[i], s.attribute_count());
        THROW(new Exception(ERROR_ATTRIBUTE_COUNT_MISMATCH, buf));
      }
      const string name = i < aliases_.size() ? aliases_[i] : s.attribute(positions_[i]).name();
      if (!b->AddAs(positions_[i], name)) THROW(Duplicate(name));
    }
    return Success(static_cast<const BoundSingleSourceProjector*>(b.release()));
  }
  virtual SingleSourceProjector* Clone() const { return new PositionProjector(positions_, aliases_); }
  virtual string ToString(bool) const { return "AttributesAt(...)"; }
 private:
  vector<int> positions_;
  vector<string> aliases_;
};

class AllProjector : public SingleSourceProjector {
 public:
  explicit AllProjector(const string& prefix) : prefix_(prefix) {}
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& s) const {
    std::unique_ptr<BoundSingleSourceProjector> b(new BoundSingleSourceProjector(s));
    for (int i = 0; i < s.attribute_count(); ++i) {
      if (!b->AddAs(i, prefix_ + s.attribute(i).name())) THROW(Duplicate(prefix_ + s.attribute(i).name()));
    }
    return Success(static_cast<const BoundSingleSourceProjector*>(b.release()));
  }
  virtual SingleSourceProjector* Clone() const { return new AllProjector(prefix_); }
  virtual string ToString(bool) const { return prefix_.empty() ? "*" : prefix_ + "*"; }
 private:
  string prefix_;
};

class RenameProjector : public SingleSourceProjector {
 public:
  RenameProjector(const vector<string>& aliases, const SingleSourceProjector* source) : aliases_(aliases), source_(source) {}
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& s) const {
    FailureOrOwned<const BoundSingleSourceProjector> inner = source_->Bind(s);
    PROPAGATE_ON_FAILURE(inner);
    if (inner->result_schema().attribute_count() != static_cast<int>(aliases_.size())) {
      THROW(new Exception(ERROR_ATTRIBUTE_COUNT_MISMATCH, "Rename: alias count differs from the projected attribute count"));
    }
    std::unique_ptr<BoundSingleSourceProjector> b(new BoundSingleSourceProjector(s));
    for (size_t i = 0; i < aliases_.size(); ++i) {
      if (!b->AddAs(inner->source_attribute_position(static_cast<int>(i)), aliases_[i])) THROW(Duplicate(aliases_[i]));
    }
    return Success(static_cast<const BoundSingleSourceProjector*>(b.release()));
  }
  virtual SingleSourceProjector* Clone() const { return new RenameProjector(aliases_, source_->Clone()); }
  virtual string ToString(bool v) const { return "RENAME(" + source_->ToString(v) + ")"; }
 private:
  vector<string> aliases_;
  std::unique_ptr<const SingleSourceProjector> source_;
};

}  // namespace

CompoundSingleSourceProjector::~CompoundSingleSourceProjector() {
  for (size_t i = 0; i < projectors_.size(); ++i) delete projectors_[i];
}
FailureOrOwned<const BoundSingleSourceProjector> CompoundSingleSourceProjector::Bind(const TupleSchema& s) const {
  std::unique_ptr<BoundSingleSourceProjector> b(new BoundSingleSourceProjector(s));
  for (size_t i = 0; i < projectors_.size(); ++i) {
    FailureOrOwned<const BoundSingleSourceProjector> inner = projectors_[i]->Bind(s);
    PROPAGATE_ON_FAILURE(inner);
    const TupleSchema& rs = inner->result_schema();
    for (int j = 0; j < rs.attribute_count(); ++j) {
      if (!b->AddAs(inner->source_attribute_position(j), rs.attribute(j).name())) THROW(Duplicate(rs.attribute(j).name()));
    }
  }
  return Success(static_cast<const BoundSingleSourceProjector*>(b.release()));
}
CompoundSingleSourceProjector* CompoundSingleSourceProjector::Clone() const {
  CompoundSingleSourceProjector* c = new CompoundSingleSourceProjector;
  for (size_t i = 0; i < projectors_.size(); ++i) c->add(projectors_[i]->Clone());
  return c;
}
string CompoundSingleSourceProjector::ToString(bool verbose) const {
  string s = "(";
  for (size_t i = 0; i < projectors_.size(); ++i) { if (i) s += ", "; s += projectors_[i]->ToString(verbose); }
  return s + ")";
}

const SingleSourceProjector* ProjectNamedAttribute(const StringPiece& name) {
  return new NamedProjector(name.as_string(), name.as_string(), false);
}
const SingleSourceProjector* ProjectNamedAttributeAs(const StringPiece& name, const StringPiece& alias) {
  return new NamedProjector(name.as_string(), alias.as_string(), true);
}
const SingleSourceProjector* ProjectAttributeAt(int position) {
  return new PositionProjector(vector<int>(1, position), vector<string>());
}
const SingleSourceProjector* ProjectAttributeAtAs(int position, const StringPiece& alias) {
  return new PositionProjector(vector<int>(1, position), vector<string>(1, alias.as_string()));
}
const SingleSourceProjector* ProjectAttributesAt(const vector<int>& positions) {
  return new PositionProjector(positions, vector<string>());
}
const SingleSourceProjector* ProjectNamedAttributes(const vector<string>& names) {
  CompoundSingleSourceProjector* c = new CompoundSingleSourceProjector;
  for (size_t i = 0; i < names.size(); ++i) c->add(ProjectNamedAttribute(names[i]));
  return c;
}
const SingleSourceProjector* ProjectAllAttributes() { return new AllProjector(""); }
const SingleSourceProjector* ProjectAllAttributes(const StringPiece& prefix) { return new AllProjector(prefix.as_string()); }
const SingleSourceProjector* ProjectRename(const vector<string>& aliases, const SingleSourceProjector* source) {
  return new RenameProjector(aliases, source);
}

// ---- multi source
BoundMultiSourceProjector::BoundMultiSourceProjector(const vector<const TupleSchema*>& source_schemas) {
  for (size_t i = 0; i < source_schemas.size(); ++i) source_schemas_.push_back(*source_schemas[i]);
}
bool BoundMultiSourceProjector::AddAs(int source_index, int attribute_position, const StringPiece& alias) {
  const Attribute& a = source_schemas_[source_index].attribute(attribute_position);
  if (!result_schema_.add_attribute(Attribute(alias.as_string(), a.type(), a.nullability()))) return false;
  sources_.push_back(source_index);
  positions_.push_back(attribute_position);
  return true;
}
CompoundMultiSourceProjector::~CompoundMultiSourceProjector() {
  for (size_t i = 0; i < projectors_.size(); ++i) delete projectors_[i].second;
}
FailureOrOwned<const BoundMultiSourceProjector> CompoundMultiSourceProjector::Bind(
    const vector<const TupleSchema*>& source_schemas) const {
  std::unique_ptr<BoundMultiSourceProjector> b(new BoundMultiSourceProjector(source_schemas));
  for (size_t i = 0; i < projectors_.size(); ++i) {
    const int src = projectors_[i].first;
    if (src < 0 || src >= static_cast<int>(source_schemas.size())) {
      THROW(new Exception(ERROR_ATTRIBUTE_COUNT_MISMATCH, "Multi-source projector refers to a source that does not exist"));
    }
    FailureOrOwned<const BoundSingleSourceProjector> inner = projectors_[i].second->Bind(*source_schemas[src]);
    PROPAGATE_ON_FAILURE(inner);
    const TupleSchema& rs = inner->result_schema();
    for (int j = 0; j < rs.attribute_count(); ++j) {
      if (!b->AddAs(src, inner->source_attribute_position(j), rs.attribute(j).name())) THROW(Duplicate(rs.attribute(j).name()));
    }
  }
  return Success(static_cast<const BoundMultiSourceProjector*>(b.release()));
}
string CompoundMultiSourceProjector::ToString(bool verbose) const {
  string s = "(";
  for (size_t i = 0; i < projectors_.size(); ++i) { if (i) s += ", "; s += projectors_[i].second->ToString(verbose); }
  return s + ")";
}

}  // namespace supersonic
