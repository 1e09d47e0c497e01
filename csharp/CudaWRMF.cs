// CudaWRMF.cs -- ItemRecommendation.WRMF (ItemRecommendation/WRMF.cs:56-180, MF.cs:37-196) on libmmlb200.so.
// Compile into MyMediaLite.dll ("CudaWRMF".CreateItemRecommender(), Extensions.cs:216-228).
// Not compiled here: no .NET/Mono toolchain in this image (INTEGRATION.md).
using System;
using System.Collections.Generic;
using System.Globalization;
using System.IO;
using System.Linq;
using MyMediaLite.Data;
using MyMediaLite.DataType;
using MyMediaLite.IO;
using MyMediaLite.Eval;
using MyMediaLite.Native;

namespace MyMediaLite.ItemRecommendation
{
	public class CudaWRMF : ItemRecommender, IIterativeModel
	{
		protected MmlHandle model, dev_feedback;
		protected readonly object gate = new object();

		public uint NumFactors { get; set; }
		public uint NumIter { get; set; }
		public double Alpha { get; set; }
		public double Regularization { get; set; }
		public double InitMean { get; set; }
		public double InitStdDev { get; set; }
		public uint NumGpus { get; set; }

		public CudaWRMF()
		{
			NumFactors = 10; NumIter = 15; Alpha = 1; Regularization = 0.015; InitStdDev = 0.1; NumGpus = 1;   // WRMF.cs:56-65, MF.cs:37-48
		}

		protected virtual void InitModel()
		{
			IntPtr ctx = Mml.Context(NumGpus), f, m;
			cache = null;
			int n = Feedback.Count;
			var users = new int[n]; var items = new int[n];
			for (int t = 0; t < n; t++) { users[t] = Feedback.Users[t]; items[t] = Feedback.Items[t]; }
			Mml.Check(Mml.mml_feedback_create(ctx, users, items, n, MaxUserID, MaxItemID, out f));
			dev_feedback = new MmlHandle(f, Mml.mml_feedback_destroy);
			var p = new MmlWrmfParams { num_factors = (int) NumFactors, alpha = Alpha, regularization = Regularization };
			Mml.Check(Mml.mml_wrmf_create(ctx, f, ref p, out m));
			model = new MmlHandle(m, Mml.mml_wrmf_destroy);
			var user_factors = new Matrix<float>(MaxUserID + 1, (int) NumFactors);   // MF.cs:51-58: user matrix first
			var item_factors = new Matrix<float>(MaxItemID + 1, (int) NumFactors);
			user_factors.InitNormal(InitMean, InitStdDev);
			item_factors.InitNormal(InitMean, InitStdDev);
			Mml.Check(Mml.mml_wrmf_set_model(m, user_factors.data, item_factors.data));
		}

		public override void Train()
		{
			lock (gate)
			{
				InitModel();
				for (uint i = 0; i < NumIter; i++) Iterate();
			}
		}

		/// <summary>WRMF.Iterate (WRMF.cs:68-73): user half-sweep, then item half-sweep</summary>
		public virtual void Iterate()
		{
			lock (gate) { Mml.Check(Mml.mml_wrmf_iterate(model.DangerousGetHandle())); cache = null; }
		}

		/// <summary>RetrainUser / RetrainItem (WRMF.cs:159-170): Gram matrix of the other side + Optimize() for the one row</summary>
		public void RetrainUser(int user_id) { lock (gate) { Mml.Check(Mml.mml_wrmf_retrain(model.DangerousGetHandle(), 0, new int[] { user_id }, 1)); cache = null; } }
		public void RetrainItem(int item_id) { lock (gate) { Mml.Check(Mml.mml_wrmf_retrain(model.DangerousGetHandle(), 1, new int[] { item_id }, 1)); cache = null; } }

		/// <summary>Eval.Items.Evaluate (Eval/Items.cs:126-209) in one device call: candidate selection (with its shuffle), the skip
		/// rules' bookkeeping and the averaging stay here, ranking and measures run on the device without materialising the lists</summary>
		public ItemRecommendationEvaluationResults Evaluate(IPosOnlyFeedback test, IPosOnlyFeedback training, IList<int> test_users = null,
			IList<int> candidate_items = null, CandidateItems candidate_item_mode = CandidateItems.OVERLAP,
			RepeatedEvents repeated_events = RepeatedEvents.No, int n = -1)
		{
			if (test_users == null) test_users = test.AllUsers;
			var cand = Items.Candidates(candidate_items, candidate_item_mode, test, training).ToArray();
			var users = test_users.ToArray();
			Func<IPosOnlyFeedback, long[]> ptr_of = fb => { var p = new long[users.Length + 1]; for (int b = 0; b < users.Length; b++) p[b + 1] = p[b] + (users[b] <= fb.MaxUserID ? fb.UserMatrix[users[b]].Count : 0); return p; };
			Func<IPosOnlyFeedback, long[], int[]> idx_of = (fb, p) => { var x = new int[Math.Max(p[users.Length], 1)]; for (int b = 0; b < users.Length; b++) if (users[b] <= fb.MaxUserID) fb.UserMatrix[users[b]].CopyTo(x, (int) p[b]); return x; };
			var test_ptr = ptr_of(test); var test_idx = idx_of(test, test_ptr);
			long[] ign_ptr = null; int[] ign_idx = null;
			if (repeated_events == RepeatedEvents.No) { ign_ptr = ptr_of(training); ign_idx = idx_of(training, ign_ptr); }
			var rows = new float[Math.Max(users.Length * 8, 1)]; var used = new int[Math.Max(users.Length, 1)];
			lock (gate) Mml.Check(Mml.mml_wrmf_evaluate(model.DangerousGetHandle(), users, users.Length, cand, cand.Length, test_ptr, test_idx, ign_ptr, ign_idx, n, rows, used));
			var result = new ItemRecommendationEvaluationResults();
			string[] names = { "AUC", "MAP", "NDCG", "MRR", "prec@5", "prec@10", "recall@5", "recall@10" };
			int num_users = 0;
			for (int b = 0; b < users.Length; b++)
			{
				if (used[b] != 1) continue;
				num_users++;
				for (int j = 0; j < 8; j++) result[names[j]] += rows[b * 8 + j];
			}
			foreach (string measure in Items.Measures) result[measure] /= num_users;
			result["num_users"] = num_users; result["num_lists"] = num_users; result["num_items"] = cand.Length;
			return result;
		}

		public override float Predict(int user_id, int item_id)
		{
			if (user_id > MaxUserID || item_id > MaxItemID) return float.MinValue;   // MF.cs:151-157
			var r = Recommend(user_id, 1, null, new int[] { item_id });
			return r.Count > 0 ? r[0].Item2 : float.MinValue;
		}

		/// <summary>all-users lists of the current model for one (n, candidates) pair, ignore rows = the training feedback</summary>
		sealed class ListCache { public int n; public int[] cand; public IList<Tuple<int, float>>[] lists; }
		ListCache cache;

		/// <summary>Recommender.Recommend (Recommender.cs:52-103) for one user. Eval.Items.Evaluate (Eval/Items.cs:147-164) and
		/// WritePredictions (ItemRecommendation/Extensions.cs:65-128) call this once per user with ignore_items = the user's training
		/// items, from TPL threads: the first such call after the model changed computes the lists of ALL users in one device call
		/// and keeps them; later calls with the same n and candidates whose ignore list is the user's training row are lookups.</summary>
		public override IList<Tuple<int, float>> Recommend(int user_id, int n = -1, ICollection<int> ignore_items = null, ICollection<int> candidate_items = null)
		{
			if (n > 0 && user_id >= 0 && user_id <= MaxUserID && ignore_items != null && Feedback != null
			    && ignore_items.Count == Feedback.UserMatrix[user_id].Count && Feedback.UserMatrix[user_id].IsSupersetOf(ignore_items))
			{
				lock (gate)
				{
					int[] cand = candidate_items == null ? Enumerable.Range(0, Math.Max(MaxItemID - 1, 0)).ToArray() : candidate_items.ToArray();
					if (cache == null || cache.n != n || !cache.cand.SequenceEqual(cand))
					{
						var all_users = Enumerable.Range(0, MaxUserID + 1).ToArray();
						var rows = new ICollection<int>[all_users.Length];
						for (int u = 0; u < rows.Length; u++) rows[u] = Feedback.UserMatrix[u];
						cache = new ListCache { n = n, cand = cand, lists = RecommendMany(all_users, n, rows, cand) };
					}
					return cache.lists[user_id];
				}
			}
			return RecommendMany(new int[] { user_id }, n, ignore_items == null ? null : new ICollection<int>[] { ignore_items }, candidate_items)[0];
		}

		/// <summary>The all-users loop of Extensions.WritePredictions / Eval.Items.Evaluate in ONE device call
		/// (tcgen05 scoring GEMM with fused top-k; exact re-scoring makes the result bit-identical to per-user Predict loops)</summary>
		public IList<Tuple<int, float>>[] RecommendMany(IList<int> users, int n, IList<ICollection<int>> ignore_items, ICollection<int> candidate_items)
		{
			int[] cand = candidate_items == null ? Enumerable.Range(0, Math.Max(MaxItemID - 1, 0)).ToArray() : candidate_items.ToArray();   // Recommender.cs:57-58
			int n_out = n < 0 ? cand.Length : Math.Min(n, cand.Length);
			var u = users.ToArray();
			long[] ptr = null; int[] idx = null;
			if (ignore_items != null)
			{
				ptr = new long[u.Length + 1];
				for (int b = 0; b < u.Length; b++) ptr[b + 1] = ptr[b] + (ignore_items[b] == null ? 0 : ignore_items[b].Count);
				idx = new int[Math.Max(ptr[u.Length], 1)];
				for (int b = 0; b < u.Length; b++) if (ignore_items[b] != null) ignore_items[b].CopyTo(idx, (int) ptr[b]);
			}
			var items = new int[Math.Max(u.Length * n_out, 1)]; var scores = new float[items.Length]; var counts = new int[u.Length];
			Mml.Check(Mml.mml_wrmf_recommend(model.DangerousGetHandle(), u, u.Length, n, cand, cand.Length, ptr, idx, items, scores, counts));
			var result = new IList<Tuple<int, float>>[u.Length];
			for (int b = 0; b < u.Length; b++)
			{
				var list = new List<Tuple<int, float>>(counts[b]);
				for (int r = 0; r < counts[b]; r++) list.Add(Tuple.Create(items[b * n_out + r], scores[b * n_out + r]));
				result[b] = list;
			}
			return result;
		}

		public override void SaveModel(string filename)
		{
			int nu = MaxUserID + 1, ni = MaxItemID + 1, k = (int) NumFactors;
			var U = new float[nu * k]; var V = new float[ni * k];
			Mml.Check(Mml.mml_wrmf_get_model(model.DangerousGetHandle(), U, V));
			using (StreamWriter writer = Model.GetWriter(filename, this.GetType(), "2.99"))   // MF.cs:160-168
			{
				writer.WriteMatrix(new Matrix<float>(nu, k) { data = U });
				writer.WriteMatrix(new Matrix<float>(ni, k) { data = V });
			}
		}

		public override string ToString()
		{
			return string.Format(CultureInfo.InvariantCulture, "{0} num_factors={1} regularization={2} alpha={3} num_iter={4}",
				this.GetType().Name, NumFactors, Regularization, Alpha, NumIter);
		}
	}
}
