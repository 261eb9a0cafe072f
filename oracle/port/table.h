// oracle/port/table.h — TEST INFRASTRUCTURE ONLY (CPU oracle).
// Shared plain types: physical column types, pushed-down filter descriptions
// (include/lingodb/runtime/storage/TableStorage.h:14-31) and the oracle's table holder, which
// replaces LingoDBTable::TableChunk (src/runtime/storage/LingoDBTable.cpp:200-225; not buildable
// here because LingoDBTable.cpp pulls the MLIR-dependent catalog).
#pragma once
#include <cstdint>
#include <string>
#include <variant>
#include <vector>

namespace oracle {

enum class FilterOp : uint8_t { EQ, NEQ, LT, LTE, GT, GTE, NOTNULL, IN }; // same order as the reference enum
enum class PhysType : uint8_t { INT32 = 0, INT64 = 1, DATE32 = 2, DECIMAL128 = 3, FSB4 = 4, STRING = 5 };

struct FilterDescription {
   std::string columnName;
   size_t columnId = 0;
   FilterOp op;
   std::variant<std::string, int64_t, double> value;
};
struct ColumnSchema {
   std::string name;
   PhysType type;
   int32_t precision = 0, scale = 0;
};

// One record batch: per column the Arrow buffers {validity, data|offsets, string bytes}.
struct HostChunk {
   int64_t numRows;
   std::vector<const void*> buffers; // 3 per column
};
struct HostTable {
   std::string name;
   std::vector<ColumnSchema> schema;
   std::vector<HostChunk> chunks;
   int64_t numRows = 0;
   int colIndex(const std::string& n) const {
      for (size_t i = 0; i < schema.size(); i++)
         if (schema[i].name == n) return (int) i;
      return -1;
   }
};

} // namespace oracle
