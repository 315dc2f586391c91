// Host-side pieces of the supersonic.h mirror that need no device: Arena, ViewCopier, Block / Table with variable-length
// cells, TableRowWriter, Limit over a host scan, ParseString* over literals + GetConstantExpressionValue, File / FileOutput /
// FileInput. Compiled against supersonic_b200/host/include and linked with libssb200_plan.so by tests/test_host_units.py.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "supersonic/supersonic.h"
#include "supersonic/cursor/core/limit.h"
#include "supersonic/cursor/infrastructure/file_io.h"

using namespace supersonic;   // NOLINT

static int failures = 0;
#define EXPECT(cond)                                                              \
  do {                                                                            \
    if (!(cond)) { ++failures; fprintf(stderr, "%s:%d: FAILED: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

static void TestArena() {
  Arena arena(16, 64);
  std::vector<StringPiece> kept;
  std::vector<std::string> want;
  for (int i = 0; i < 200; ++i) {
    std::string s(static_cast<size_t>(i % 37), static_cast<char>('a' + i % 26));
    const char* p = arena.AddStringPieceContent(s);
    EXPECT(p != NULL || s.empty());
    kept.push_back(StringPiece(p, s.size()));
    want.push_back(s);
  }
  for (size_t i = 0; i < kept.size(); ++i) EXPECT(kept[i] == StringPiece(want[i]));   // earlier pieces stay valid as the arena grows
  EXPECT(arena.memory_footprint() >= 3000);
  arena.Reset();
  EXPECT(arena.memory_footprint() == 0);
  MemoryLimit tiny(8);
  Arena limited(&tiny, 16, 64);
  EXPECT(limited.AddStringPieceContent("more than eight bytes") == NULL);   // the allocator refuses
}

static void TestBlockViewCopierTable() {
  TupleSchema schema;
  schema.add_attribute(Attribute("id", INT32, NOT_NULLABLE));
  schema.add_attribute(Attribute("name", STRING, NULLABLE));
  Table table(schema, HeapBufferAllocator::Get());
  TableRowWriter writer(&table);
  {
    std::string volatile_name = "Terry";
    writer.AddRow().Int32(1).String(volatile_name).CheckSuccess();
    volatile_name = "XXXXX";   // the table owns a copy
  }
  writer.AddRow().Int32(2).Null().CheckSuccess();
  writer.AddRow().Int32(3).String("").CheckSuccess();
  EXPECT(table.row_count() == 3);
  const View& v = table.view();
  EXPECT(v.column(1).typed_data<STRING>()[0] == StringPiece("Terry"));
  EXPECT(v.column(1).is_null()[1] && !v.column(1).is_null()[0] && !v.column(1).is_null()[2]);
  // deep copy into a block at an offset; the source goes away
  Block block(schema, HeapBufferAllocator::Get());
  EXPECT(block.Reallocate(8));
  EXPECT(block.view().row_count() == 8);   // a block's view spans its capacity (block.h:467)
  {
    Table temp(schema, HeapBufferAllocator::Get());
    EXPECT(temp.AppendView(v) == 3);
    ViewCopier copier(schema, /* deep copy */ true);
    EXPECT(copier.Copy(3, temp.view(), 2, &block) == 3);
    EXPECT(copier.Copy(7, temp.view(), 2, &block) == 0);   // does not fit: nothing copied
  }
  EXPECT(block.view().column(0).typed_data<INT32>()[2] == 1 && block.view().column(0).typed_data<INT32>()[4] == 3);
  EXPECT(block.view().column(1).typed_data<STRING>()[2] == StringPiece("Terry"));
  EXPECT(block.view().column(1).is_null()[3]);
}

static void TestLimitOverHostScan() {
  TupleSchema schema;
  schema.add_attribute(Attribute("a", INT64, NOT_NULLABLE));
  std::vector<int64> data(1000);
  for (size_t i = 0; i < data.size(); ++i) data[i] = static_cast<int64>(i) * 3;
  View view(schema);
  view.set_row_count(data.size());
  view.mutable_column(0)->Reset(data.data(), NULL);
  std::unique_ptr<Operation> op(Limit(10, 25, ScanView(view)));
  std::unique_ptr<Cursor> cursor(SucceedOrDie(op->CreateCursor()));
  int64 expect = 30, rows = 0;
  for (;;) {
    ResultView r = cursor->Next(7);
    if (r.is_eos()) break;
    EXPECT(r.has_data());
    if (!r.has_data()) break;
    EXPECT(r.view().row_count() <= 7);
    for (rowcount_t i = 0; i < r.view().row_count(); ++i, ++rows, expect += 3) EXPECT(r.view().column(0).typed_data<INT64>()[i] == expect);
  }
  EXPECT(rows == 25);
}

static void TestParseConstants() {
  bool is_null = true;
  std::unique_ptr<const Expression> date(ParseStringNulling(DATE, ConstString(" 1991/01/01 ")));
  FailureOr<int32> d = GetConstantExpressionValue<DATE>(*date, &is_null);
  EXPECT(d.is_success() && !is_null && d.get() == 7670);
  std::unique_ptr<const Expression> bad(ParseStringNulling(DATE, ConstString("Mort")));
  FailureOr<int32> b = GetConstantExpressionValue<DATE>(*bad, &is_null);
  EXPECT(b.is_success() && is_null);
  std::unique_ptr<const Expression> num(ParseStringQuiet(INT64, ConstString("-42")));
  FailureOr<int64> n = GetConstantExpressionValue<INT64>(*num, &is_null);
  EXPECT(n.is_success() && !is_null && n.get() == -42);
  std::unique_ptr<const Expression> str(ConstString("abc"));
  FailureOr<std::string> s = GetConstantExpressionValue<STRING>(*str, &is_null);
  EXPECT(s.is_success() && !is_null && s.get() == "abc");
  std::unique_ptr<const Expression> wrong(ConstInt32(5));
  FailureOr<int64> w = GetConstantExpressionValue<INT64>(*wrong, &is_null);
  EXPECT(w.is_failure() && w.exception().return_code() == ERROR_ATTRIBUTE_TYPE_MISMATCH);
  std::unique_ptr<const Expression> column(ParseStringNulling(INT32, NamedAttribute("s")));
  TupleSchema schema;
  schema.add_attribute(Attribute("s", STRING, NOT_NULLABLE));
  FailureOrOwned<BoundExpressionTree> bound = column->Bind(schema, HeapBufferAllocator::Get(), 16);
  EXPECT(bound.is_failure() && bound.exception().return_code() == ERROR_NOT_IMPLEMENTED);   // parsing a column: refused, not guessed
}

static void TestFileRoundTrip(const char* path) {
  TupleSchema schema;
  schema.add_attribute(Attribute("id", INT32, NOT_NULLABLE));
  schema.add_attribute(Attribute("name", STRING, NULLABLE));
  Table table(schema, HeapBufferAllocator::Get());
  TableRowWriter writer(&table);
  for (int i = 0; i < 20000; ++i) {
    writer.AddRow().Int32(i);
    if (i % 11 == 0) writer.Null(); else writer.String(std::string(static_cast<size_t>(i % 5), 'x') + "y");
  }
  writer.CheckSuccess();
  std::unique_ptr<Sink> sink(FileOutput(File::OpenOrDie(path, "w"), TAKE_OWNERSHIP));
  FailureOrVoid written = WriteCursor(SucceedOrDie(table.CreateCursor()), sink.get());
  EXPECT(written.is_success());
  EXPECT(sink->Finalize().is_success());
  std::unique_ptr<Cursor> scan(SucceedOrDie(FileInput(schema, File::OpenOrDie(path, "r"), /* delete when done */ true, HeapBufferAllocator::Get())));
  int rows = 0;
  for (;;) {
    ResultView r = scan->Next(1000);
    if (r.is_eos()) break;
    EXPECT(r.has_data());
    if (!r.has_data()) break;
    for (rowcount_t i = 0; i < r.view().row_count(); ++i, ++rows) {
      EXPECT(r.view().column(0).typed_data<INT32>()[i] == rows);
      const bool isn = r.view().column(1).is_null()[i];
      EXPECT(isn == (rows % 11 == 0));
      if (!isn) EXPECT(r.view().column(1).typed_data<STRING>()[i] == StringPiece(std::string(static_cast<size_t>(rows % 5), 'x') + "y"));
    }
  }
  EXPECT(rows == 20000);
  scan.reset();
  EXPECT(!File::Exists(path));   // delete_when_done
}

int main(int argc, char** argv) {
  TestArena();
  TestBlockViewCopierTable();
  TestLimitOverHostScan();
  TestParseConstants();
  TestFileRoundTrip(argc > 1 ? argv[1] : "/tmp/ssb200_host_units.ssb");
  printf(failures == 0 ? "OK host units\n" : "%d FAILED\n", failures);
  return failures == 0 ? 0 : 1;
}
