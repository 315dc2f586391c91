// supersonic/projector.h -- column projections (base/infrastructure/projector.h:71-422).
#ifndef SUPERSONIC_B200_HOST_PROJECTOR_H_
#define SUPERSONIC_B200_HOST_PROJECTOR_H_

#include "supersonic/base.h"

namespace supersonic {

// A bound projection of one source: result column i = source column source_position(i),
// under the name and type given by result_schema().
class BoundSingleSourceProjector {
 public:
  explicit BoundSingleSourceProjector(const TupleSchema& source_schema) : source_schema_(source_schema) {}
  const TupleSchema& source_schema() const { return source_schema_; }
  const TupleSchema& result_schema() const { return result_schema_; }
  int source_attribute_position(int result_position) const { return positions_[result_position]; }
  bool Add(int source_position) { return AddAs(source_position, source_schema_.attribute(source_position).name()); }
  bool AddAs(int source_position, const StringPiece& alias);
  // result view = the projected columns of `source` (pointer shuffling only, project.cc:49-59)
  void Project(const View& source, View* target) const;
  bool IsAttributeProjected(int source_position) const;
 private:
  TupleSchema source_schema_, result_schema_;
  vector<int> positions_;
};

class SingleSourceProjector {
 public:
  virtual ~SingleSourceProjector() {}
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& source_schema) const = 0;
  virtual SingleSourceProjector* Clone() const = 0;
  virtual string ToString(bool verbose) const = 0;
 protected:
  SingleSourceProjector() {}
};

class CompoundSingleSourceProjector : public SingleSourceProjector {
 public:
  CompoundSingleSourceProjector() {}
  virtual ~CompoundSingleSourceProjector();
  // takes ownership
  CompoundSingleSourceProjector* add(const SingleSourceProjector* projector) { projectors_.push_back(projector); return this; }
  virtual FailureOrOwned<const BoundSingleSourceProjector> Bind(const TupleSchema& source_schema) const;
  virtual CompoundSingleSourceProjector* Clone() const;
  virtual string ToString(bool verbose) const;
 private:
  vector<const SingleSourceProjector*> projectors_;
};

const SingleSourceProjector* ProjectNamedAttribute(const StringPiece& name);
const SingleSourceProjector* ProjectNamedAttributeAs(const StringPiece& name, const StringPiece& alias);
const SingleSourceProjector* ProjectAttributeAt(int position);
const SingleSourceProjector* ProjectAttributeAtAs(int position, const StringPiece& alias);
const SingleSourceProjector* ProjectAttributesAt(const vector<int>& positions);
const SingleSourceProjector* ProjectNamedAttributes(const vector<string>& names);
const SingleSourceProjector* ProjectAllAttributes();
const SingleSourceProjector* ProjectAllAttributes(const StringPiece& prefix);
const SingleSourceProjector* ProjectRename(const vector<string>& aliases, const SingleSourceProjector* source);

// Projection over several sources (hash join result): result column i = column
// source_attribute_position(i) of source source_index(i).
class BoundMultiSourceProjector {
 public:
  explicit BoundMultiSourceProjector(const vector<const TupleSchema*>& source_schemas);
  const TupleSchema& result_schema() const { return result_schema_; }
  int source_count() const { return static_cast<int>(source_schemas_.size()); }
  const TupleSchema& source_schema(int i) const { return source_schemas_[i]; }
  int source_index(int result_position) const { return sources_[result_position]; }
  int source_attribute_position(int result_position) const { return positions_[result_position]; }
  bool AddAs(int source_index, int attribute_position, const StringPiece& alias);
 private:
  vector<TupleSchema> source_schemas_;
  TupleSchema result_schema_;
  vector<int> sources_, positions_;
};

class MultiSourceProjector {
 public:
  virtual ~MultiSourceProjector() {}
  virtual FailureOrOwned<const BoundMultiSourceProjector> Bind(const vector<const TupleSchema*>& source_schemas) const = 0;
  virtual string ToString(bool verbose) const = 0;
 protected:
  MultiSourceProjector() {}
};

class CompoundMultiSourceProjector : public MultiSourceProjector {
 public:
  CompoundMultiSourceProjector() {}
  virtual ~CompoundMultiSourceProjector();
  // takes ownership of the projector
  CompoundMultiSourceProjector* add(int source_index, const SingleSourceProjector* projector) {
    projectors_.push_back(std::make_pair(source_index, projector));
    return this;
  }
  virtual FailureOrOwned<const BoundMultiSourceProjector> Bind(const vector<const TupleSchema*>& source_schemas) const;
  virtual string ToString(bool verbose) const;
 private:
  vector<std::pair<int, const SingleSourceProjector*> > projectors_;
};

}  // namespace supersonic
#endif  // SUPERSONIC_B200_HOST_PROJECTOR_H_
