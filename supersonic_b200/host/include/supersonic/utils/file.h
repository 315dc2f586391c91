// supersonic/utils/file.h:36-106: the File abstraction the reference's file cursors read and write through (global
// class, as in the reference). One implementation: a local file over stdio. A File is closed with Close(), which
// also deletes the object (file.cc: "delete this").
#ifndef SUPERSONIC_B200_HOST_UTILS_FILE_H_
#define SUPERSONIC_B200_HOST_UTILS_FILE_H_
#include <stdio.h>

#include <string>

#include "supersonic/base.h"

class File {
 public:
  // Create() does not open the file; OpenOrDie() aborts when it cannot.
  static File* Create(const std::string& file_name, const std::string& mode);
  static File* OpenOrDie(const std::string& file_name, const std::string& mode);
  static bool Exists(const std::string& file);
  static std::string JoinPath(const std::string& dirname, const std::string& basename);
  virtual bool Exists() const;
  virtual bool Open();
  virtual bool Delete();
  virtual bool Close();
  virtual int64 Read(void* buffer, uint64 length);
  virtual char* ReadLine(char* buffer, uint64 max_length);
  virtual int64 Write(const void* buffer, uint64 length);
  virtual bool Seek(int64 position);
  virtual bool eof();
  virtual const std::string& CreateFileName() { return create_file_name_; }
 protected:
  File(const std::string& create_file_name, const std::string& mode) : create_file_name_(create_file_name), mode_(mode), f_(NULL) {}
  virtual ~File();
 private:
  std::string create_file_name_, mode_;
  FILE* f_;
};
#endif
