// CudaBiasedMatrixFactorization.cs / CudaMatrixFactorization -- the reference's MatrixFactorization and
// BiasedMatrixFactorization surface (RatingPrediction/MatrixFactorization.cs:35-418,
// RatingPrediction/BiasedMatrixFactorization.cs:61-563) on libmmlb200.so. Same property names and defaults,
// same ToString() shape, same model-file layout; one added property, NumGpus.
// Compile into MyMediaLite.dll so that "CudaBiasedMatrixFactorization".CreateRatingPredictor() (Extensions.cs:170-182)
// finds it. Not compiled here: no .NET/Mono toolchain in this image (INTEGRATION.md).
using System;
using System.Collections.Generic;
using System.Globalization;
using System.IO;
using System.Linq;
using MyMediaLite.Data;
using MyMediaLite.DataType;
using MyMediaLite.Eval;
using MyMediaLite.IO;
using MyMediaLite.Native;

namespace MyMediaLite.RatingPrediction
{
	/// <summary>MatrixFactorization on the GPU: r = global_bias + p_u . q_i, SGD (MatrixFactorization.cs:166-196)</summary>
	public class CudaMatrixFactorization : RatingPredictor, IIterativeModel, IFoldInRatingPredictor
	{
		/// <summary>the device model; shared by clones, replaced (never mutated in place) by Train()</summary>
		protected MmlHandle model, dev_ratings;
		protected readonly object gate = new object();

		public float InitMean { get; set; }
		public float InitStdDev { get; set; }
		public uint NumFactors { get; set; }
		public float LearnRate { get; set; }
		public float Decay { get; set; }
		public virtual float Regularization { get; set; }
		public uint NumIter { get; set; }
		/// <summary>GPUs of one host to train on (engine property; default 1)</summary>
		public uint NumGpus { get; set; }

		protected virtual bool Biased { get { return false; } }

		public CudaMatrixFactorization()
		{
			// MatrixFactorization.cs:87-96
			Regularization = 0.015f; LearnRate = 0.01f; Decay = 1.0f; NumIter = 30; InitStdDev = 0.1f; NumFactors = 10; NumGpus = 1;
		}

		/// <summary>IList -> int[]/float[] of length Count (StaticRatings exposes its arrays, Ratings/RatingsProxy do not)</summary>
		protected static T[] AsArray<T>(IList<T> list, int count)
		{
			var a = list as T[];
			if (a != null && a.Length == count) return a;
			var r = new T[count];
			for (int i = 0; i < count; i++) r[i] = list[i];
			return r;
		}

		protected virtual MmlMfParams Params()
		{
			MmlMfParams p;
			Mml.mml_mf_params_default(out p);
			p.biased = Biased ? 1 : 0;
			p.num_factors = (int) NumFactors; p.learn_rate = LearnRate; p.decay = Decay; p.regularization = Regularization;
			p.schedule = Parallel ? Mml.SCHEDULE_DSGD : Mml.SCHEDULE_SERIAL;
			return p;
		}

		/// <summary>threads the reference would use: 1 for plain MF, MaxThreads for the biased model</summary>
		protected virtual int Threads { get { return 1; } }

		/// <summary>whether Iterate() runs the parallel epoch kernel (Mml.Order); several GPUs always do</summary>
		protected bool Parallel
		{
			get
			{
				if (NumGpus > 1 || Mml.Order == Mml.EngineOrder.Parallel) return true;
				if (Mml.Order == Mml.EngineOrder.Reference) return Threads > 1;
				return Threads > 1 || (ratings != null && ratings.Count >= Mml.SerialBelow);
			}
		}
		bool parallel_model;

		/// <summary>InitModel (MatrixFactorization.cs:99-116): draws come from MyMediaLite.Random in the reference's order</summary>
		protected internal virtual void InitModel()
		{
			IntPtr ctx = Mml.Context(NumGpus), r, m;
			int n = ratings.Count;
			Mml.Check(Mml.mml_ratings_create(ctx, AsArray(ratings.Users, n), AsArray(ratings.Items, n), AsArray(ratings.Values, n), n, MaxUserID, MaxItemID, out r));
			dev_ratings = new MmlHandle(r, Mml.mml_ratings_destroy);
			var p = Params();
			parallel_model = p.schedule == Mml.SCHEDULE_DSGD;
			Mml.Check(Mml.mml_sgd_create(ctx, r, ref p, null, null, out m));
			model = new MmlHandle(m, Mml.mml_sgd_destroy);
			var user_factors = new Matrix<float>(MaxUserID + 1, (int) NumFactors);
			var item_factors = new Matrix<float>(MaxItemID + 1, (int) NumFactors);
			user_factors.InitNormal(InitMean, InitStdDev);
			item_factors.InitNormal(InitMean, InitStdDev);
			Mml.Check(Mml.mml_sgd_set_model(m, user_factors.data, item_factors.data, null, null));   // rows without ratings are zeroed by the library
		}

		public override void Train()
		{
			lock (gate)
			{
				InitModel();
				for (uint it = 0; it < NumIter; it++) Iterate();
			}
		}

		public virtual void Iterate()
		{
			lock (gate)
			{
				if (parallel_model)
				{
					int G, W; long rounds, staged;
					Mml.Check(Mml.mml_sgd_strata_info(model.DangerousGetHandle(), out G, out W, out rounds, out staged));
					var subepoch_sequence = new List<int>(Enumerable.Range(0, G));
					subepoch_sequence.Shuffle();                            // BiasedMatrixFactorization.cs:210-211
					Mml.Check(Mml.mml_sgd_iterate(model.DangerousGetHandle(), subepoch_sequence.ToArray(), null, 0));
					return;
				}
				var index = AsArray(ratings.RandomIndex, ratings.Count);
				Mml.Check(Mml.mml_sgd_iterate(model.DangerousGetHandle(), null, index, index.Length));
			}
		}

		public override float Predict(int user_id, int item_id)
		{
			var res = new float[1];
			Mml.Check(Mml.mml_sgd_predict(model.DangerousGetHandle(), new int[] { user_id }, new int[] { item_id }, 1, res));
			return res[0];
		}

		/// <summary>Eval.Ratings.Evaluate (Eval/Ratings.cs:96-139) in one device pass: RMSE, MAE, NMAE, CBD</summary>
		public Dictionary<string, float> Evaluate(IRatings test)
		{
			int n = test.Count;
			var r = new float[4];
			Mml.Check(Mml.mml_sgd_evaluate(model.DangerousGetHandle(), AsArray(test.Users, n), AsArray(test.Items, n), AsArray(test.Values, n), n, r));
			return new Dictionary<string, float> { { "RMSE", r[0] }, { "MAE", r[1] }, { "NMAE", r[2] }, { "CBD", r[3] } };
		}

		public virtual float ComputeObjective()
		{
			double v;
			Mml.Check(Mml.mml_sgd_objective(model.DangerousGetHandle(), out v));
			return (float) v;
		}

		/// <summary>length of a fold-in vector: the biased model puts the user bias in front (BiasedMatrixFactorization.cs:80-82)</summary>
		protected int FoldInStride { get { return (int) NumFactors + (Biased ? 1 : 0); } }

		/// <summary>RetrainUser (MatrixFactorization.cs:141-149, BiasedMatrixFactorization.cs:419-424): re-draw the row, zero the bias,
		/// LearnFactors (:198-202) = NumIter passes over ByUser[user_id] updating the user side only</summary>
		public virtual void RetrainUser(int user_id) { Retrain(user_id, false, ratings.ByUser[user_id]); }

		/// <summary>RetrainItem (MatrixFactorization.cs:152-160, BiasedMatrixFactorization.cs:426-431)</summary>
		public virtual void RetrainItem(int item_id) { Retrain(item_id, true, ratings.ByItem[item_id]); }

		void Retrain(int id, bool by_item, IList<int> indices)
		{
			lock (gate)
			{
				var row = new float[NumFactors];
				row.InitNormal(InitMean, InitStdDev);
				Mml.Check(Mml.mml_sgd_set_rows(model.DangerousGetHandle(), by_item ? 1 : 0, new int[] { id }, 1, row, Biased ? new float[] { 0 } : null));
				var idx = indices.ToArray();
				// LearnFactors (MatrixFactorization.cs:198-202): NumIter passes over the list
				Mml.Check(Mml.mml_sgd_learn_factors(model.DangerousGetHandle(), idx, idx.Length, by_item ? 0 : 1, by_item ? 1 : 0, (int) NumIter));
			}
		}

		/// <summary>FoldIn (MatrixFactorization.cs:323-347, BiasedMatrixFactorization.cs:445-492): the vector and the shuffle are drawn
		/// here (same RNG order as the reference), the SGD passes run on the device</summary>
		protected virtual float[] FoldIn(IList<Tuple<int, float>> rated_items)
		{
			var init = new float[NumFactors];
			init.InitNormal(InitMean, InitStdDev);
			rated_items.Shuffle();
			var items = rated_items.Select(t => t.Item1).ToArray();
			var values = rated_items.Select(t => t.Item2).ToArray();
			var result = new float[FoldInStride];
			Mml.Check(Mml.mml_sgd_fold_in(model.DangerousGetHandle(), new long[] { 0, items.Length }, items, values, 1, init, (int) NumIter, result));
			return result;
		}

		/// <summary>ScoreItems (MatrixFactorization.cs:350-363)</summary>
		public IList<Tuple<int, float>> ScoreItems(IList<Tuple<int, float>> rated_items, IList<int> candidate_items)
		{
			var user_vector = FoldIn(rated_items);
			var cand = candidate_items.ToArray();
			var scores = new float[cand.Length];
			Mml.Check(Mml.mml_sgd_score_items(model.DangerousGetHandle(), user_vector, 1, cand, cand.Length, scores));
			var result = new Tuple<int, float>[cand.Length];
			for (int i = 0; i < cand.Length; i++) result[i] = Tuple.Create(cand[i], scores[i]);
			return result;
		}

		/// <summary>LoadModel (MatrixFactorization.cs:386-408): the factors go to a fresh device model</summary>
		public override void LoadModel(string filename)
		{
			using (StreamReader reader = Model.GetReader(filename, this.GetType()))
			{
				var bias = float.Parse(reader.ReadLine(), CultureInfo.InvariantCulture);
				var user_factors = (Matrix<float>) reader.ReadMatrix(new Matrix<float>(0, 0));
				var item_factors = (Matrix<float>) reader.ReadMatrix(new Matrix<float>(0, 0));
				if (user_factors.NumberOfColumns != item_factors.NumberOfColumns)
					throw new IOException(string.Format("Number of user and item factors must match: {0} != {1}", user_factors.NumberOfColumns, item_factors.NumberOfColumns));
				Adopt(user_factors, item_factors, null, null, bias, MinRating, MaxRating);
			}
		}

		/// <summary>a model without training data: one pseudo rating per id keeps every row (InitModel zeroes rows without ratings)</summary>
		protected void Adopt(Matrix<float> U, Matrix<float> V, float[] bu, float[] bi, float bias, float min, float max)
		{
			lock (gate)
			{
				MaxUserID = U.NumberOfRows - 1; MaxItemID = V.NumberOfRows - 1; NumFactors = (uint) U.NumberOfColumns;
				int n = Math.Max(U.NumberOfRows, V.NumberOfRows);
				var uu = new int[n]; var ii = new int[n]; var vv = new float[n];
				for (int t = 0; t < n; t++) { uu[t] = t % U.NumberOfRows; ii[t] = t % V.NumberOfRows; vv[t] = min; }
				IntPtr ctx = Mml.Context(), r, m;
				Mml.Check(Mml.mml_ratings_create(ctx, uu, ii, vv, n, MaxUserID, MaxItemID, out r));
				dev_ratings = new MmlHandle(r, Mml.mml_ratings_destroy);
				var p = Params(); p.schedule = Mml.SCHEDULE_SERIAL;
				Mml.Check(Mml.mml_sgd_create(ctx, r, ref p, null, null, out m));
				model = new MmlHandle(m, Mml.mml_sgd_destroy);
				Mml.Check(Mml.mml_sgd_set_model(m, U.data, V.data, bu, bi));
				Mml.Check(Mml.mml_sgd_set_scale(m, min, max, bias));
			}
		}

		public override void SaveModel(string filename)
		{
			var U = new float[(MaxUserID + 1) * NumFactors]; var V = new float[(MaxItemID + 1) * NumFactors];
			float gb, lr;
			Mml.Check(Mml.mml_sgd_get_model(model.DangerousGetHandle(), U, V, null, null, out gb, out lr));
			using (StreamWriter writer = Model.GetWriter(filename, this.GetType(), "2.99"))
			{
				writer.WriteLine(gb.ToString(CultureInfo.InvariantCulture));
				writer.WriteMatrix(new Matrix<float>(MaxUserID + 1, (int) NumFactors) { data = U });
				writer.WriteMatrix(new Matrix<float>(MaxItemID + 1, (int) NumFactors) { data = V });
			}
		}

		public override string ToString()
		{
			return string.Format(CultureInfo.InvariantCulture,
				"{0} num_factors={1} regularization={2} learn_rate={3} learn_rate_decay={4} num_iter={5}",
				this.GetType().Name, NumFactors, Regularization, LearnRate, Decay, NumIter);
		}
	}

	/// <summary>BiasedMatrixFactorization on the GPU (BiasedMatrixFactorization.cs:61-563)</summary>
	public class CudaBiasedMatrixFactorization : CudaMatrixFactorization
	{
		public float BiasReg { get; set; }
		public float BiasLearnRate { get; set; }
		public float RegU { get; set; }
		public float RegI { get; set; }
		public override float Regularization { set { base.Regularization = value; RegU = value; RegI = value; } }   // :97-104
		public bool FrequencyRegularization { get; set; }
		public OptimizationTarget Loss { get; set; }
		public int MaxThreads { get; set; }
		public bool BoldDriver { get; set; }
		public bool NaiveParallelization { get; set; }

		protected override bool Biased { get { return true; } }

		public CudaBiasedMatrixFactorization() : base()
		{
			BiasReg = 0.01f; BiasLearnRate = 1.0f; MaxThreads = 1;   // :85-141
		}

		protected override MmlMfParams Params()
		{
			var p = base.Params();
			p.bias_learn_rate = BiasLearnRate; p.bias_reg = BiasReg; p.reg_u = RegU; p.reg_i = RegI;
			p.frequency_regularization = FrequencyRegularization ? 1 : 0;
			p.loss = Loss == OptimizationTarget.MAE ? Mml.LOSS_MAE : (Loss == OptimizationTarget.LogisticLoss ? Mml.LOSS_LOGISTIC : Mml.LOSS_RMSE);
			p.bold_driver = BoldDriver ? 1 : 0; p.max_threads = MaxThreads;
			// MaxThreads > 1 selects the reference's DSGD block schedule (:178-184); on the GPU the worker groups are CTAs, and the
			// parallel kernel is also what MaxThreads = 1 runs unless the engine order says Reference (base.Params, Mml.Order).
			// NaiveParallelization (:136-141, :201-204): the list schedule of MultiCore.PartitionIndices (one GPU only)
			if (p.schedule == Mml.SCHEDULE_DSGD && NaiveParallelization && NumGpus <= 1) p.schedule = Mml.SCHEDULE_NAIVE;
			return p;
		}

		protected override int Threads { get { return MaxThreads; } }

		public override void SaveModel(string filename)
		{
			int nu = MaxUserID + 1, ni = MaxItemID + 1, k = (int) NumFactors;
			var U = new float[nu * k]; var V = new float[ni * k]; var bu = new float[nu]; var bi = new float[ni];
			float gb, lr;
			Mml.Check(Mml.mml_sgd_get_model(model.DangerousGetHandle(), U, V, bu, bi, out gb, out lr));
			using (StreamWriter writer = Model.GetWriter(filename, this.GetType(), "2.99"))   // layout of :339-351
			{
				writer.WriteLine(gb.ToString(CultureInfo.InvariantCulture));
				writer.WriteLine(min_rating.ToString(CultureInfo.InvariantCulture));
				writer.WriteLine(max_rating.ToString(CultureInfo.InvariantCulture));
				writer.WriteVector(bu);
				writer.WriteMatrix(new Matrix<float>(nu, k) { data = U });
				writer.WriteVector(bi);
				writer.WriteMatrix(new Matrix<float>(ni, k) { data = V });
			}
		}

		/// <summary>LoadModel (BiasedMatrixFactorization.cs:353-402)</summary>
		public override void LoadModel(string filename)
		{
			using (StreamReader reader = Model.GetReader(filename, this.GetType()))
			{
				var bias = float.Parse(reader.ReadLine(), CultureInfo.InvariantCulture);
				var min = float.Parse(reader.ReadLine(), CultureInfo.InvariantCulture);
				var max = float.Parse(reader.ReadLine(), CultureInfo.InvariantCulture);
				var bu = reader.ReadVector();
				var user_factors = (Matrix<float>) reader.ReadMatrix(new Matrix<float>(0, 0));
				var bi = reader.ReadVector();
				var item_factors = (Matrix<float>) reader.ReadMatrix(new Matrix<float>(0, 0));
				if (user_factors.NumberOfColumns != item_factors.NumberOfColumns)
					throw new IOException(string.Format("Number of user and item factors must match: {0} != {1}", user_factors.NumberOfColumns, item_factors.NumberOfColumns));
				if (bu.Count != user_factors.dim1 || bi.Count != item_factors.dim1)
					throw new IOException("Number of biases must match the number of factor rows");
				min_rating = min; max_rating = max;
				Adopt(user_factors, item_factors, bu.ToArray(), bi.ToArray(), bias, min, max);
			}
		}

		public override string ToString()
		{
			return string.Format(CultureInfo.InvariantCulture,
				"{0} num_factors={1} bias_reg={2} reg_u={3} reg_i={4} frequency_regularization={5} learn_rate={6} bias_learn_rate={7} learn_rate_decay={8} num_iter={9} bold_driver={10} loss={11} max_threads={12} naive_parallelization={13}",
				this.GetType().Name, NumFactors, BiasReg, RegU, RegI, FrequencyRegularization, LearnRate, BiasLearnRate, Decay, NumIter, BoldDriver, Loss, MaxThreads, NaiveParallelization);
		}
	}
}
