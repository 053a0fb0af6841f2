/*
 * BenchmarkB200Config2 — the reference's own CPU path for BASELINE.json configs[1], as a JMH benchmark (BASELINE.md §4,
 * baseline A).  SOURCE ONLY: the build image has no JDK, Maven or Pinot jars, so this class has never been compiled or run
 * here; it is kept ready for a box that has them (drop it into pinot-perf/src/main/java/org/apache/pinot/perf/ of a Pinot
 * 1.4.0-SNAPSHOT checkout and run it like the other pinot-perf benchmarks).
 *
 * What it measures: the server side of ONE query -- InstancePlanMakerImplV2.makeInstancePlan(...).execute(), i.e. the real
 * GroupByCombineOperator over 8 mmap'd immutable segments of 12.5 M rows each (100 M rows, the 8 columns the query touches),
 * inverted index on c1 skipped so that both predicates scan -- on a thread pool of ALL host cores with
 * maxExecutionThreads raised accordingly (the default caps a query at min(10, cores / 2) threads,
 * QueryMultiThreadingUtils.java:46-47).  rows/s = 100e6 / average time.  The broker reduce and the DataTable serialisation
 * are NOT included: bench.py's `value` stops at the merged per-server result too.
 *
 * The rows come from a generator with the column shapes of pinot_b200/datagen.py (uniform dictIds, table-wide dimension
 * dictionaries, cardinalities 1000 / 10000 / 8 / 16 / 32 / 100000 x 3, metric values in [0, 10^6)); java.util.SplittableRandom
 * cannot reproduce numpy's PCG64 streams, so the tables are statistically identical, not byte-identical.
 */
package org.apache.pinot.perf;

import java.io.File;
import java.util.ArrayList;
import java.util.List;
import java.util.SplittableRandom;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.TimeUnit;
import java.util.stream.Collectors;
import org.apache.commons.io.FileUtils;
import org.apache.pinot.core.operator.blocks.InstanceResponseBlock;
import org.apache.pinot.core.plan.Plan;
import org.apache.pinot.core.plan.maker.InstancePlanMakerImplV2;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.core.query.request.context.utils.QueryContextConverterUtils;
import org.apache.pinot.segment.local.indexsegment.immutable.ImmutableSegmentLoader;
import org.apache.pinot.segment.local.segment.creator.impl.SegmentIndexCreationDriverImpl;
import org.apache.pinot.segment.local.segment.index.loader.IndexLoadingConfig;
import org.apache.pinot.segment.local.segment.readers.GenericRowRecordReader;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.segment.spi.SegmentContext;
import org.apache.pinot.segment.spi.creator.SegmentGeneratorConfig;
import org.apache.pinot.spi.config.table.TableConfig;
import org.apache.pinot.spi.config.table.TableType;
import org.apache.pinot.spi.data.FieldSpec;
import org.apache.pinot.spi.data.Schema;
import org.apache.pinot.spi.data.readers.GenericRow;
import org.apache.pinot.spi.utils.builder.TableConfigBuilder;
import org.openjdk.jmh.annotations.Benchmark;
import org.openjdk.jmh.annotations.BenchmarkMode;
import org.openjdk.jmh.annotations.Fork;
import org.openjdk.jmh.annotations.Measurement;
import org.openjdk.jmh.annotations.Mode;
import org.openjdk.jmh.annotations.OutputTimeUnit;
import org.openjdk.jmh.annotations.Param;
import org.openjdk.jmh.annotations.Scope;
import org.openjdk.jmh.annotations.Setup;
import org.openjdk.jmh.annotations.State;
import org.openjdk.jmh.annotations.TearDown;
import org.openjdk.jmh.annotations.Warmup;
import org.openjdk.jmh.runner.Runner;
import org.openjdk.jmh.runner.options.OptionsBuilder;


@BenchmarkMode(Mode.AverageTime)
@OutputTimeUnit(TimeUnit.MILLISECONDS)
@Fork(value = 1, jvmArgs = {"-Xmx32g", "-XX:MaxDirectMemorySize=16g"})
@Warmup(iterations = 5, time = 5)
@Measurement(iterations = 5, time = 5)
@State(Scope.Benchmark)
public class BenchmarkB200Config2 {
  private static final File INDEX_DIR = new File(FileUtils.getTempDirectory(), "BenchmarkB200Config2");
  private static final String TABLE = "t";
  private static final String[] DIMS = {"c1", "c2", "d0", "d1", "d2"};
  private static final int[] DIM_CARDS = {1000, 10000, 8, 16, 32};
  private static final String[] METRICS = {"m0", "m1", "m2"};
  private static final int METRIC_CARD = 100_000;

  @Param({"8"})
  int _numSegments;
  @Param({"12500000"})
  int _rowsPerSegment;
  @Param({"16", "500"})          // 16 values ~ 0.8 % of the rows pass the filter, 500 values ~ 25 %
  int _inValues;

  private final List<IndexSegment> _segments = new ArrayList<>();
  private ExecutorService _executor;
  private String _query;
  private int[][] _dimDictionaries;

  public static void main(String[] args)
      throws Exception {
    new Runner(new OptionsBuilder().include(BenchmarkB200Config2.class.getSimpleName()).build()).run();
  }

  @Setup
  public void setUp()
      throws Exception {
    FileUtils.deleteQuietly(INDEX_DIR);
    Schema.SchemaBuilder sb = new Schema.SchemaBuilder().setSchemaName(TABLE);
    for (String d : DIMS) {
      sb.addSingleValueDimension(d, FieldSpec.DataType.INT);
    }
    for (String m : METRICS) {
      sb.addMetric(m, FieldSpec.DataType.INT);
    }
    Schema schema = sb.build();
    TableConfig tableConfig =
        new TableConfigBuilder(TableType.OFFLINE).setTableName(TABLE).setInvertedIndexColumns(List.of("c1", "d0")).build();

    // table-wide dimension dictionaries: `card` distinct values out of [0, 10 * card)
    SplittableRandom dictRandom = new SplittableRandom(42);
    _dimDictionaries = new int[DIMS.length][];
    for (int j = 0; j < DIMS.length; j++) {
      _dimDictionaries[j] = dictRandom.ints(0, Math.max(DIM_CARDS[j] * 10, 1000)).distinct().limit(DIM_CARDS[j]).sorted().toArray();
    }
    for (int s = 0; s < _numSegments; s++) {
      buildSegment(schema, tableConfig, s);
      _segments.add(ImmutableSegmentLoader.load(new File(INDEX_DIR, "seg_" + s), new IndexLoadingConfig(tableConfig, schema)));
    }

    int cores = Runtime.getRuntime().availableProcessors();
    _executor = Executors.newFixedThreadPool(cores);
    int[] c1 = _dimDictionaries[0];
    int step = Math.max(1, c1.length / _inValues);
    List<String> in = new ArrayList<>();
    for (int i = 0; i < c1.length && in.size() < _inValues; i += step) {
      in.add(Integer.toString(c1[i]));
    }
    int k = _dimDictionaries[1][_dimDictionaries[1].length / 2];
    _query = "SET maxExecutionThreads = " + cores + "; SET skipIndexes = 'c1=inverted'; "
        + "SELECT d0, d1, d2, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t WHERE c1 IN (" + String.join(", ", in) + ") AND c2 < " + k
        + " GROUP BY d0, d1, d2 LIMIT 100000";
  }

  private void buildSegment(Schema schema, TableConfig tableConfig, int index)
      throws Exception {
    SplittableRandom random = new SplittableRandom(42L + index);
    // per-segment metric dictionaries: 100 000 distinct values out of [0, 10^6)
    int[][] metricDictionaries = new int[METRICS.length][];
    for (int j = 0; j < METRICS.length; j++) {
      metricDictionaries[j] = random.ints(0, 1_000_000).distinct().limit(METRIC_CARD).toArray();
    }
    List<GenericRow> rows = new ArrayList<>(_rowsPerSegment);
    for (int i = 0; i < _rowsPerSegment; i++) {
      GenericRow row = new GenericRow();
      for (int j = 0; j < DIMS.length; j++) {
        row.putValue(DIMS[j], _dimDictionaries[j][random.nextInt(DIM_CARDS[j])]);
      }
      for (int j = 0; j < METRICS.length; j++) {
        row.putValue(METRICS[j], metricDictionaries[j][random.nextInt(METRIC_CARD)]);
      }
      rows.add(row);
    }
    SegmentGeneratorConfig config = new SegmentGeneratorConfig(tableConfig, schema);
    config.setOutDir(INDEX_DIR.getPath());
    config.setTableName(TABLE);
    config.setSegmentName("seg_" + index);
    SegmentIndexCreationDriverImpl driver = new SegmentIndexCreationDriverImpl();
    driver.init(config, new GenericRowRecordReader(rows));
    driver.build();
  }

  @TearDown
  public void tearDown() {
    for (IndexSegment segment : _segments) {
      segment.destroy();
    }
    _executor.shutdownNow();
    FileUtils.deleteQuietly(INDEX_DIR);
  }

  /** one server-side execution of the query: plan + GroupByCombineOperator over all segments + the merged result block */
  @Benchmark
  public InstanceResponseBlock query()
      throws Exception {
    QueryContext queryContext = QueryContextConverterUtils.getQueryContext(_query);
    queryContext.setEndTimeMs(System.currentTimeMillis() + 600_000L);
    List<SegmentContext> contexts = _segments.stream().map(SegmentContext::new).collect(Collectors.toList());
    Plan plan = new InstancePlanMakerImplV2().makeInstancePlan(contexts, queryContext, _executor, null);
    return plan.execute();
  }
}
