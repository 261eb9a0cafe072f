// oracle/port/pipelines.h — TEST INFRASTRUCTURE ONLY (CPU oracle).  Query entry points of pipelines.cpp.
#pragma once
#include "table.h"
#include "values.h"

#include <string>
#include <vector>

namespace oracle {

struct Q6Params {
   std::string shipdateGe, shipdateLt, discountGe, discountLe;
   int64_t quantityLt;
};
struct Q6Result {
   i128 revenue; // decimal(24,4)
   double seconds;
};
Q6Result runQ6(const HostTable& lineitem, const Q6Params& p);

struct Q1Params {
   std::string shipdateLe;
};
struct Q1Row {
   int32_t returnflag, linestatus;
   int64_t sumQty, sumBasePrice;  // decimal(12,2)
   i128 sumDiscPrice;             // decimal(33,4)
   i128 sumCharge;                // decimal(38,6)
   i128 avgQty, avgPrice, avgDisc; // decimal(31,21)
   int64_t count;
};
std::vector<Q1Row> runQ1(const HostTable& lineitem, const Q1Params& p, double* seconds);

struct Q3Params {
   std::string segment, date;
};
struct Q3Row {
   int32_t orderkey;
   i128 revenue; // decimal(33,4)
   int32_t orderdate; // date32
   int32_t shippriority;
};
std::vector<Q3Row> runQ3(const HostTable& customer, const HostTable& orders, const HostTable& lineitem, const Q3Params& p, double* seconds);

struct Q5Params {
   std::string regionName, dateGe, dateLt;
};
struct Q5Row {
   std::string name;
   i128 revenue; // decimal(33,4)
};
std::vector<Q5Row> runQ5(const HostTable& customer, const HostTable& orders, const HostTable& lineitem, const HostTable& supplier, const HostTable& nation, const HostTable& region, const Q5Params& p, double* seconds);

struct Q9Params {
   std::string needle; // p_name like '%<needle>%'
};
struct Q9Row {
   std::string nation;
   int64_t year;
   i128 sumProfit; // decimal(38,4) raw
};
std::vector<Q9Row> runQ9(const HostTable& part, const HostTable& supplier, const HostTable& lineitem, const HostTable& partsupp, const HostTable& orders, const HostTable& nation, const Q9Params& p, double* seconds);

// Q4 / Q12: oracle twins prepared for the next widening step (semi-join with marker, conditional sums); no GPU operator yet
struct Q4Params {
   std::string dateGe, dateLt;
};
struct Q4Row {
   std::string priority;
   int64_t orderCount;
};
std::vector<Q4Row> runQ4(const HostTable& orders, const HostTable& lineitem, const Q4Params& p, double* seconds);
struct Q12Params {
   std::string mode1, mode2, dateGe, dateLt;
};
struct Q12Row {
   std::string shipmode;
   int64_t highLineCount, lowLineCount;
};
std::vector<Q12Row> runQ12(const HostTable& orders, const HostTable& lineitem, const Q12Params& p, double* seconds);

// Q18: a group-by with as many groups as orders (the PreAggregationHashtable merge path at scale), HAVING, semi-join, top-100
struct Q18Row {
   std::string name;
   int32_t custkey, orderkey, orderdate;
   int64_t totalprice, sumQuantity; // decimal(12,2) raw
};
std::vector<Q18Row> runQ18(const HostTable& customer, const HostTable& orders, const HostTable& lineitem, int64_t quantityGt, double* seconds);

} // namespace oracle
