// The helpers of supersonic/utils/map_util.h that client code of the reference uses (test/guide/*.cc):
// FindOrNull, FindWithDefault, ContainsKey, InsertIfNotPresent, InsertOrUpdate. Written from their contracts.
#ifndef SUPERSONIC_B200_HOST_UTILS_MAP_UTIL_H_
#define SUPERSONIC_B200_HOST_UTILS_MAP_UTIL_H_
#include <utility>

template <class Collection>
const typename Collection::value_type::second_type* FindOrNull(const Collection& c,
                                                               const typename Collection::value_type::first_type& key) {
  typename Collection::const_iterator it = c.find(key);
  return it == c.end() ? 0 : &it->second;
}
template <class Collection>
typename Collection::value_type::second_type* FindOrNull(Collection& c,   // NOLINT
                                                         const typename Collection::value_type::first_type& key) {
  typename Collection::iterator it = c.find(key);
  return it == c.end() ? 0 : &it->second;
}
template <class Collection>
const typename Collection::value_type::second_type& FindWithDefault(const Collection& c,
                                                                    const typename Collection::value_type::first_type& key,
                                                                    const typename Collection::value_type::second_type& value) {
  typename Collection::const_iterator it = c.find(key);
  return it == c.end() ? value : it->second;
}
template <class Collection, class Key>
bool ContainsKey(const Collection& c, const Key& key) { return c.find(key) != c.end(); }
template <class Collection>
bool InsertIfNotPresent(Collection* c, const typename Collection::value_type::first_type& key,
                        const typename Collection::value_type::second_type& value) {
  return c->insert(typename Collection::value_type(key, value)).second;
}
template <class Collection>
bool InsertOrUpdate(Collection* c, const typename Collection::value_type::first_type& key,
                    const typename Collection::value_type::second_type& value) {
  std::pair<typename Collection::iterator, bool> r = c->insert(typename Collection::value_type(key, value));
  if (!r.second) r.first->second = value;
  return r.second;
}
#endif
